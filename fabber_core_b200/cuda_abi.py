"""ctypes mirror of include/fabber_cuda.h (the thin host<->device C ABI under Vb::DoCalculations).

The structs here must stay byte-compatible with the header; tests/test_abi.py checks sizeof()
against the values the compiled library reports.
"""
import ctypes as C
import os

import numpy as np

MAX_PARAMS = 8
MAX_PHIS = 4
AR_NOISE_FIELDS = 7
AR_CROSS_BY_NAME = {"none": 0, "same": 1, "dual": 2}


def ar2_noise_fields(n_alphas):
    """FABBER_CUDA_AR2_NOISE_FIELDS: b1 c1 b2 c2, alpha means, packed alpha precisions"""
    return 4 + n_alphas + n_alphas * (n_alphas + 1) // 2


OK, ERR_INVALID, ERR_CUDA, ERR_BAD_VOXEL = 0, -1, -2, -3
MODEL_LINEAR, MODEL_POLY, MODEL_EXP = 1, 2, 3
MODEL_PLUGIN = 100
MODEL_ORACLE_SINE = 101   # known to the test oracle only (oracle/vb_oracle.cc), never to the device library
NOISE_WHITE, NOISE_AR1 = 0, 1
METHOD_VB, METHOD_NLLS = 0, 1
CONV_MAXITS, CONV_FCHANGE, CONV_FREDUCE, CONV_TRIALMODE, CONV_LM = 0, 1, 2, 3, 4
CONV_BY_NAME = {
    "maxits": CONV_MAXITS,
    "pointzeroone": CONV_FCHANGE,
    "freduce": CONV_FREDUCE,
    "trialmode": CONV_TRIALMODE,
    "lm": CONV_LM,
}
VOX_SETUP_FLAG = 0x100


class Model(C.Structure):
    _fields_ = [
        ("id", C.c_int),
        ("n_params", C.c_int),
        ("design", C.c_void_p),
        ("poly_degree", C.c_int),
        ("exp_num", C.c_int),
        ("exp_dt", C.c_double),
        ("plugin_launchers", C.c_void_p),
        ("consts", C.c_double * 16),
        ("design_len", C.c_int),
    ]


class Param(C.Structure):
    _fields_ = [
        ("transform", C.c_char),
        ("prior_type", C.c_char),
        ("pad_", C.c_char * 6),
        ("prior_mean", C.c_double),
        ("prior_prec", C.c_double),
        ("prior_var", C.c_double),
        ("post_mean", C.c_double),
        ("post_var", C.c_double),
    ]


class VbProblem(C.Structure):
    _fields_ = [
        ("n_voxels", C.c_int),
        ("n_times", C.c_int),
        ("model", Model),
        ("params", Param * MAX_PARAMS),
        ("noise_type", C.c_int),
        ("n_phis", C.c_int),
        ("phi_pattern", C.c_void_p),
        ("time_masked", C.c_void_p),
        ("noise_prior_b", C.c_double * MAX_PHIS),
        ("noise_prior_c", C.c_double * MAX_PHIS),
        ("noise_post_b", C.c_double * MAX_PHIS),
        ("noise_post_c", C.c_double * MAX_PHIS),
        ("locked_noise_stdev", C.c_double),
        ("ar_alpha_prior_prec", C.c_double),
        ("ar_cross_terms", C.c_int),
        ("conv_type", C.c_int),
        ("max_iterations", C.c_int),
        ("fchange", C.c_double),
        ("max_trials", C.c_int),
        ("need_f", C.c_int),
        ("f_history_len", C.c_int),
        ("allow_bad_voxels", C.c_int),
        ("spatial_dims", C.c_int),
        ("spatial_speed", C.c_double),
        ("spatial_q1", C.c_double),
        ("spatial_q2", C.c_double),
        ("update_first_iter", C.c_int),
        ("nx", C.c_int),
        ("ny", C.c_int),
        ("nz", C.c_int),
        ("method", C.c_int),
        ("nlls_lm", C.c_int),
        ("nlls_have_start", C.c_int),
        ("nlls_start", C.c_double * MAX_PARAMS),
    ]


class VbBuffers(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("image_prior", C.c_void_p * MAX_PARAMS),
        ("init_mean", C.c_void_p),
        ("init_cov", C.c_void_p),
        ("init_noise", C.c_void_p),
        ("lock_centre", C.c_void_p),
        ("coords", C.c_void_p),
        ("mean", C.c_void_p),
        ("cov", C.c_void_p),
        ("noise", C.c_void_p),
        ("free_energy", C.c_void_p),
        ("f_history", C.c_void_p),
        ("iterations", C.c_void_p),
        ("status", C.c_void_p),
        ("spatial_ak", C.c_void_p),
    ]


class SlabPart(C.Structure):
    """mirror of fabber_cuda_slab_part (fabber_cuda_vb_spatial_multi)"""
    _fields_ = [
        ("device", C.c_int),
        ("v0", C.c_int), ("v1", C.c_int), ("own0", C.c_int), ("own1", C.c_int),
        ("own_z0", C.c_int), ("own_z1", C.c_int),
        ("buf", VbBuffers),
    ]


# ------------------------------------------------------------------------------------------------
# Parameter transforms, host side (transforms.h:114-242, transforms.cc:17-25) - used to build the
# Fabber-space prior exactly as FwdModel::GetParameters does (fwdmodel.cc:268-277).
# ------------------------------------------------------------------------------------------------
def _to_model(code, v):
    if code == "L":
        return np.exp(v)
    if code == "S":
        return np.log(1 + np.exp(v)) if v < 10 else v
    if code == "F":
        return 1 / (1 + np.exp(v))
    if code == "A":
        return abs(v)
    return v


def _to_fabber(code, v):
    with np.errstate(all="ignore"):
        if code == "L":
            return float(np.log(v))
        if code == "S":
            return float(np.log(np.exp(v) - 1)) if v < 10 else v
        if code == "F":
            return float(np.log(1 / v - 1))
    return v


def _to_fabber_var(code, v):
    with np.errstate(all="ignore"):
        if code == "L":
            return float(np.log(v))
        if code in ("I", "F"):
            return v
        return float(_to_fabber(code, _to_model(code, 0.0) + np.sqrt(v)) ** 2)


class ProblemSpec(object):
    """Python-side description of one VB problem; owns the numpy arrays the C structs point at."""

    def __init__(self, model, n_times, params=None, design=None, degree=None, num_exps=1, dt=1.0,
                 noise="white", noise_pattern="1", masked_timepoints=(), prior_noise_stddev=None,
                 locked_noise_stdev=-1.0, convergence="maxits", max_iterations=10, fchange=0.01,
                 max_trials=10, need_f=None, f_history_len=0, allow_bad_voxels=False,
                 prior_types=None, spatial_dims=3, spatial_speed=-1.0, spatial_q1=10.0, spatial_q2=1.0,
                 update_first_iter=False, param_overrides=None, plugin_launchers=None, num_echoes=1,
                 ar_cross_terms="none", method="vb", nlls_lm=False, nlls_start=None):
        self.keep = []
        self.n_times = int(n_times)
        m = Model()
        defaults = []  # (name, prior_mean, prior_var, post_mean, post_var, transform) model space
        if model == "linear":
            design = np.ascontiguousarray(design, dtype=np.float64)
            assert design.shape[0] == n_times
            self.keep.append(design)
            m.id, m.n_params, m.design = MODEL_LINEAR, design.shape[1], design.ctypes.data
            defaults = [("Parameter_%d" % (i + 1), 0.0, 1e12, 0.0, 1e12, "I") for i in range(design.shape[1])]
        elif model == "poly":
            m.id, m.n_params, m.poly_degree = MODEL_POLY, degree + 1, degree
            defaults = [("c%d" % i, 0.0, 1e12, 0.0, 1e12, "I") for i in range(degree + 1)]
        elif model == "exp":
            m.id, m.n_params, m.exp_num, m.exp_dt = MODEL_EXP, 2 * num_exps, num_exps, dt
            for i in range(num_exps):
                defaults.append(("amp%d" % (i + 1), 1.0, 1e5, 1.0, 1.5, "L"))
                defaults.append(("r%d" % (i + 1), 1.0, 1e5, 1.0, 1.5, "L"))
        elif model == "sine":
            # the example plug-in model (fabber_core_b200/examples/sine_model.cu): a*sin(b*(t-c))+d.
            # plugin_launchers=None describes it to the TEST ORACLE (which knows it as model id 101);
            # with the table of a loaded plug-in library it is a device problem.
            m.n_params = 4
            m.consts[0] = dt
            m.id = MODEL_PLUGIN if plugin_launchers else MODEL_ORACLE_SINE
            m.plugin_launchers = plugin_launchers
            defaults = [("a", 1.0, 1e6, 1.0, 1e6, "I"), ("b", 1.0, 1e6, 1.0, 1e6, "I"), ("c", 0.0, 1e6, 0.0, 1e6, "I"),
                        ("d", 0.0, 1e6, 0.0, 1e6, "I")]
        else:
            raise ValueError("no device Evaluate hook for model %r" % model)
        self.param_names = [d[0] for d in defaults]
        P = m.n_params
        if P > MAX_PARAMS:
            raise ValueError("too many parameters")
        prob = VbProblem()
        prob.n_times = self.n_times
        prob.model = m
        types = list(prior_types) if prior_types else ["N"] * P
        param_overrides = param_overrides or {}
        for i, (name, pm, pv, qm, qv, tr) in enumerate(defaults):
            ov = param_overrides.get(name, {})
            tr = ov.get("transform", tr)
            pm = ov.get("mean", pm)
            if "prec" in ov:
                pv = 1.0 / ov["prec"]
            if 1.0 / pv > 1e12:  # fwdmodel.cc:268-271
                pv = 1e-12
            p = prob.params[i]
            p.transform = tr.encode()
            p.prior_type = ov.get("type", types[i]).encode()
            fm = _to_fabber(tr, pm)
            fv = _to_fabber_var(tr, pv)
            p.prior_mean, p.prior_var = fm, fv
            with np.errstate(all="ignore"):
                p.prior_prec = float(np.float64(1.0) / np.float64(fv))
            p.post_mean, p.post_var = qm, qv
        # noise
        if noise == "white":
            prob.noise_type = NOISE_WHITE
            pat = []
            for ch in noise_pattern:
                if "1" <= ch <= "9":
                    pat.append(ord(ch) - ord("0"))
                elif ch.isalpha():
                    pat.append(ord(ch.lower()) - ord("a") + 10)
                else:
                    raise ValueError("bad noise pattern")
            nphi = max(pat)
            full = [pat[i % len(pat)] - 1 for i in range(n_times)]
            pat_arr = np.array(full, dtype=np.uint8)
            self.keep.append(pat_arr)
            prob.n_phis = nphi
            prob.phi_pattern = pat_arr.ctypes.data
            for i in range(nphi):
                if prior_noise_stddev is None:
                    prob.noise_prior_b[i], prob.noise_prior_c[i] = 1e6, 1e-6
                    prob.noise_post_b[i], prob.noise_post_c[i] = 1e-8, 50.0
                else:
                    c = 0.5
                    b = 1 / (prior_noise_stddev * prior_noise_stddev * c)
                    prob.noise_prior_b[i] = prob.noise_post_b[i] = b
                    prob.noise_prior_c[i] = prob.noise_post_c[i] = c
        elif noise == "ar":
            prob.noise_type = NOISE_AR1
            if num_echoes not in (1, 2) or (num_echoes == 1 and ar_cross_terms != "none"):
                raise ValueError("num_echoes 1 or 2; cross terms need two echoes")
            prob.n_phis = num_echoes
            prob.ar_cross_terms = AR_CROSS_BY_NAME[ar_cross_terms]
            for i in range(num_echoes):
                prob.noise_prior_b[i], prob.noise_prior_c[i] = 1e6, 1e-6
                prob.noise_post_b[i], prob.noise_post_c[i] = 1e-8, 1e-6
            prob.ar_alpha_prior_prec = 1e-4
        else:
            raise ValueError(noise)
        if len(masked_timepoints):
            mt = np.zeros(n_times, dtype=np.uint8)
            for t in masked_timepoints:  # 1-based as in --mt<n>
                mt[t - 1] = 1
            self.keep.append(mt)
            prob.time_masked = mt.ctypes.data
        prob.locked_noise_stdev = locked_noise_stdev
        prob.conv_type = CONV_BY_NAME[convergence]
        prob.max_iterations = max_iterations
        prob.fchange = fchange
        prob.max_trials = max_trials
        uses_f = convergence != "maxits"
        prob.need_f = int(uses_f if need_f is None else (need_f or uses_f))
        prob.f_history_len = f_history_len
        prob.allow_bad_voxels = int(allow_bad_voxels)
        prob.spatial_dims = spatial_dims
        prob.spatial_speed = spatial_speed
        prob.spatial_q1, prob.spatial_q2 = spatial_q1, spatial_q2
        prob.update_first_iter = int(update_first_iter)
        prob.method = {"vb": METHOD_VB, "nlls": METHOD_NLLS}[method]
        prob.nlls_lm = int(nlls_lm)
        if nlls_start is not None:
            prob.nlls_have_start = 1
            for i, x in enumerate(nlls_start):
                prob.nlls_start[i] = x
        self.prob = prob
        self.P = P
        self.n_alphas = 2 + prob.ar_cross_terms if noise == "ar" else 0
        if noise == "ar":
            self.NN = AR_NOISE_FIELDS if num_echoes == 1 else ar2_noise_fields(self.n_alphas)
        else:
            self.NN = 2 * prob.n_phis
        self.ncov = P * (P + 1) // 2


def library_path():
    if os.environ.get("FABBER_CUDA_LIB"):  # developer knob: kernel-tuning builds (csrc/Makefile SUBSET=1)
        return os.environ["FABBER_CUDA_LIB"]
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(here, "csrc", "libfabber_cuda.so")
