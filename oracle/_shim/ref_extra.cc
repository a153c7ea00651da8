/*
 * ref_extra.cc - TEST INFRASTRUCTURE ONLY, linked into oracle/_ref/libfabbercore_ref.so next to the
 * reference's unchanged sources. The reference registers its example "exp" model through a separately
 * loaded model library (examples/exp_models.cc, fabber_load_models); fabber_destroy also tears the model
 * factory down (fabber_capi.cc:279). This adds one C entry point to (re-)register "exp" after fabber_new.
 */
#include "fabber_core/fwdmodel.h"
#include "examples/fwdmodel_exp.h"

extern "C" int fabber_ref_register_exp(void)
{
    FwdModelFactory::GetInstance()->Add("exp", &ExpFwdModel::NewInstance);
    return 0;
}

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void fabber_ref_segv(int sig)
{
    void *frames[64];
    int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    _exit(139);
}
/* debugging aid for the test-only reference build: print a native backtrace on SIGSEGV */
extern "C" void fabber_ref_install_segv_handler(void) { signal(SIGSEGV, fabber_ref_segv); }

/* Test-only accessor: the reference keeps its results as doubles (rundata.h:628) but its C API narrows
 * them to float32 (rundata_array.cc:68-98). This hands the doubles out unchanged, [rows][nvoxels] over
 * the masked voxels, so the oracle can be pinned far below float32 precision. Returns rows, <0 on error. */
#include "fabber_core/rundata_array.h"
extern "C" int fabber_ref_get_data_double(void *fab, const char *name, double *buf, int max_values)
{
    try
    {
        FabberRunDataArray *rundata = (FabberRunDataArray *)fab;
        const NEWMAT::Matrix &m = rundata->GetVoxelData(name);
        if (m.Nrows() * m.Ncols() > max_values)
            return -2;
        for (int r = 1; r <= m.Nrows(); r++)
            for (int c = 1; c <= m.Ncols(); c++)
                buf[(size_t)(r - 1) * m.Ncols() + (c - 1)] = m(r, c);
        return m.Nrows();
    }
    catch (...)
    {
        return -1;
    }
}
