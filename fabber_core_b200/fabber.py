"""ctypes binding over libfabbercore_b200.so - the same calls, in the same order and with the same array
conventions, as the reference's own Python wrapper (py/fabber.py:489-772, a Python 2 module): option
strings via fabber_set_opt (booleans = empty value), data as Fortran-flattened float32 arrays
[x, y, z(, t)], mask as int32, outputs read back with fabber_get_data_size / fabber_get_data.

Plumbing only: every number is produced by the C++ host + CUDA libraries; a missing library raises.
"""
import ctypes as C
import os

import numpy as np

FABBER_ERR_FATAL = -255


class FabberException(RuntimeError):
    """py/fabber.py:204-213"""

    def __init__(self, msg, errcode=None, log=None):
        self.errcode = errcode
        self.log = log
        RuntimeError.__init__(self, "%s (code %s)" % (msg, errcode) if errcode is not None else msg)


class FabberRun(object):
    """py/fabber.py:462-487: output data by name + the log text"""

    def __init__(self, data, log):
        self.data = data
        self.log = log


def library_path():
    here = os.path.dirname(os.path.abspath(__file__))
    return os.environ.get("FABBERCORE_B200_LIB", os.path.join(here, "libfabbercore_b200.so"))


class Fabber(object):
    def __init__(self, lib=None):
        path = lib or library_path()
        if not os.path.exists(path):
            raise FabberException("host library %s is missing: run `python -c 'import __graft_entry__ as g; "
                                  "g.build()'`" % path)
        self.clib = C.CDLL(path)
        self.errbuf = C.create_string_buffer(255)
        self.outbuf = C.create_string_buffer(1000000)
        self.progress_cb_type = C.CFUNCTYPE(None, C.c_int, C.c_int)
        c = self.clib
        c_int_arr = np.ctypeslib.ndpointer(dtype=np.int32, ndim=1, flags="CONTIGUOUS")
        c_float_arr = np.ctypeslib.ndpointer(dtype=np.float32, ndim=1, flags="CONTIGUOUS")
        # py/fabber.py:725-764
        c.fabber_new.argtypes = [C.c_char_p]
        c.fabber_new.restype = C.c_void_p
        c.fabber_load_models.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        c.fabber_set_extent.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, c_int_arr, C.c_char_p]
        c.fabber_set_opt.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
        c.fabber_set_data.argtypes = [C.c_void_p, C.c_char_p, C.c_uint, c_float_arr, C.c_char_p]
        c.fabber_get_data_size.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        c.fabber_get_data.argtypes = [C.c_void_p, C.c_char_p, c_float_arr, C.c_char_p]
        c.fabber_dorun.argtypes = [C.c_void_p, C.c_uint, C.c_char_p, C.c_char_p, self.progress_cb_type]
        c.fabber_destroy.argtypes = [C.c_void_p]
        c.fabber_destroy.restype = None
        c.fabber_get_options.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint, C.c_char_p, C.c_char_p]
        for fn in (c.fabber_get_models, c.fabber_get_methods, c.fabber_get_model_params,
                   c.fabber_get_model_param_descs, c.fabber_get_model_outputs):
            fn.argtypes = [C.c_void_p, C.c_uint, C.c_char_p, C.c_char_p]
        c.fabber_model_evaluate.argtypes = [C.c_void_p, C.c_uint, c_float_arr, C.c_uint, c_float_arr, c_float_arr,
                                            C.c_char_p]
        self.handle = None
        self._new_handle()

    # ---- handle management -----------------------------------------------------------------------
    def _new_handle(self):
        self._destroy_handle()
        self.handle = self.clib.fabber_new(self.errbuf)
        if not self.handle:
            raise FabberException("Error creating fabber context (%s)" % self.errbuf.value.decode())

    def _destroy_handle(self):
        if getattr(self, "handle", None):
            self.clib.fabber_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def _trycall(self, call, *args):
        ret = call(*args)
        if ret < 0:
            raise FabberException(self.errbuf.value.decode(errors="replace"), ret, self.outbuf.value.decode(errors="replace"))
        return ret

    def _set_options(self, rundata):
        for key, value in rundata.items():
            if key == "loadmodels":   # not an option of the run: load the library (py/fabber.py:507-513)
                self.load_models(value)
                continue
            if isinstance(value, bool):  # boolean options: key present with empty value
                if not value:
                    continue
                value = ""
            self._trycall(self.clib.fabber_set_opt, self.handle, str(key).encode(), str(value).encode(), self.errbuf)

    def load_models(self, libpath):
        """register the models of a plug-in library (fabber_load_models; py/fabber.py:505-513)"""
        self._trycall(self.clib.fabber_load_models, self.handle, str(libpath).encode(), self.errbuf)

    # ---- self description ------------------------------------------------------------------------
    def get_methods(self):
        self._trycall(self.clib.fabber_get_methods, self.handle, len(self.outbuf), self.outbuf, self.errbuf)
        return self.outbuf.value.decode().splitlines()

    def get_models(self):
        self._trycall(self.clib.fabber_get_models, self.handle, len(self.outbuf), self.outbuf, self.errbuf)
        return self.outbuf.value.decode().splitlines()

    def get_options(self, method=None, model=None):
        if method:
            key, value = b"method", method.encode()
        elif model:
            key, value = b"model", model.encode()
        else:
            key, value = None, None
        self._trycall(self.clib.fabber_get_options, self.handle, key, value, len(self.outbuf), self.outbuf, self.errbuf)
        lines = self.outbuf.value.decode().split("\n")
        opts = []
        for line in lines[1:]:
            f = line.split("\t")
            if len(f) >= 5:
                opts.append({"name": f[0], "description": f[1], "type": f[2], "optional": f[3] == "1", "default": f[4]})
        return opts, lines[0]

    def get_model_params(self, rundata):
        self._new_handle()
        self._set_options(rundata)
        self._trycall(self.clib.fabber_get_model_params, self.handle, len(self.outbuf), self.outbuf, self.errbuf)
        return self.outbuf.value.decode().splitlines()

    def model_evaluate(self, rundata, params, nt, indata=None):
        self._new_handle()
        self._set_options(rundata)
        plist = np.ascontiguousarray(params, dtype=np.float32)
        ret = np.zeros(nt, dtype=np.float32)
        indata = np.zeros(nt, dtype=np.float32) if indata is None else np.ascontiguousarray(indata, dtype=np.float32)
        self._trycall(self.clib.fabber_model_evaluate, self.handle, len(plist), plist, nt, indata, ret, self.errbuf)
        return ret

    # ---- run (py/fabber.py:634-713) -----------------------------------------------------------------
    def run_with_data(self, rundata, data, mask=None, progress_cb=None, extra_outputs=()):
        """data: dict name -> array [x, y, z] or [x, y, z, t]; mask: [x, y, z] or None."""
        if "data" not in data:
            raise FabberException("Main voxel data not provided")
        s = data["data"].shape
        nv = s[0] * s[1] * s[2]
        if mask is None:
            mask = np.ones(nv)
        mask = np.ascontiguousarray(np.asarray(mask).flatten(order="F"), dtype=np.int32)
        self._new_handle()
        self._set_options(rundata)
        self._trycall(self.clib.fabber_get_model_params, self.handle, len(self.outbuf), self.outbuf, self.errbuf)
        params = self.outbuf.value.decode().splitlines()
        output_items = []
        for opt, prefix in (("save-mean", "mean_"), ("save-std", "std_"), ("save-zstat", "zstat_"), ("save-var", "var_")):
            if opt in rundata:
                output_items += [prefix + p for p in params]
        for opt, name in (("save-noise-mean", "noise_means"), ("save-noise-std", "noise_stdevs"),
                          ("save-free-energy", "freeEnergy"), ("save-model-fit", "modelfit"),
                          ("save-residuals", "residuals"), ("save-mvn", "finalMVN"),
                          ("save-free-energy-history", "freeEnergyHistory")):
            if opt in rundata:
                output_items.append(name)
        output_items += list(extra_outputs)
        self._trycall(self.clib.fabber_set_extent, self.handle, s[0], s[1], s[2], mask, self.errbuf)
        for key, item in data.items():
            size = 1 if item.ndim == 3 else item.shape[3]
            flat = np.ascontiguousarray(np.asarray(item).flatten(order="F"), dtype=np.float32)
            self._trycall(self.clib.fabber_set_data, self.handle, key.encode(), size, flat, self.errbuf)
        cb = self.progress_cb_type(progress_cb) if progress_cb is not None else self.progress_cb_type(0)
        self._trycall(self.clib.fabber_dorun, self.handle, len(self.outbuf), self.outbuf, self.errbuf, cb)
        log = self.outbuf.value.decode(errors="replace")
        retdata = {}
        for key in output_items:
            size = self._trycall(self.clib.fabber_get_data_size, self.handle, key.encode(), self.errbuf)
            arr = np.empty(nv * size, dtype=np.float32)
            self._trycall(self.clib.fabber_get_data, self.handle, key.encode(), arr, self.errbuf)
            retdata[key] = arr.reshape([s[0], s[1], s[2], size] if size > 1 else [s[0], s[1], s[2]], order="F")
        return FabberRun(retdata, log)


FabberLib = Fabber
