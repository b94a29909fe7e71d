"""ctypes binding of libidrk.so (include/idrk.h).  Fails loudly: no fallback of any kind."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDRK_LIB") or os.path.join(_HERE, "csrc", "libidrk.so")      # IDRK_LIB: A/B runs of two builds

MAX_LEVELS = 32
HASH_REFERENCE = 0
HASH_TRILINEAR = 1
HASH_NGP = 2          # tiny-cuda-nn grid semantics (include/idrk.h)


class IdrkError(RuntimeError):
    pass


class HashGridDesc(ctypes.Structure):
    """Mirror of idrk_hashgrid_t."""
    _fields_ = [
        ("n_levels", ctypes.c_int32),
        ("n_feat", ctypes.c_int32),
        ("frac_mode", ctypes.c_int32),
        ("n_fourier", ctypes.c_int32),
        ("res", ctypes.c_float * MAX_LEVELS),
        ("rows", ctypes.c_uint32 * MAX_LEVELS),
        ("tables", ctypes.c_void_p * MAX_LEVELS),
        ("fourier_B", ctypes.c_void_p),
    ]


_lib = None
LAUNCHES = [0]          # number of kernel launches issued through the C ABI (bench.py reports it)


class _Profile:
    """Optional per-entry-point device timing with CUDA events (bench.py's roofline leg)."""

    def __init__(self):
        self.enabled = False
        self.records = []
        self.pending_flops = None
        self.pending_tag = None
        self.detail = False

    def reset(self, enabled, detail=False):
        self.enabled = enabled
        self.records = []
        self.pending_flops = None
        self.pending_tag = None
        self.detail = detail

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e, flops in self.records:
            d = out.setdefault(name, {"ms": 0.0, "flops": 0.0, "calls": 0})
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            if flops is not None:
                d["flops"] += float(flops()) if callable(flops) else float(flops)
        return out


PROFILE = _Profile()


class _Proxy:
    """Counts launches and (when profiling) brackets every call with CUDA events on the current stream."""

    def __init__(self, cdll):
        self._cdll = cdll
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._cdll, name)

            def fn(*args, _raw=raw, _name=name):
                if PROFILE.enabled:
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    flops, PROFILE.pending_flops = PROFILE.pending_flops, None
                    tag, PROFILE.pending_tag = PROFILE.pending_tag, None
                    s.record()
                    rc = _raw(*args)
                    e.record()
                    if PROFILE.detail:
                        key = _name + (tag or "")
                    else:
                        t = tag or ""
                        key = _name + "_2cta" if t.startswith("_2cta") else (_name + "_big" if t.startswith("_big") else _name)
                    PROFILE.records.append((key, s, e, flops))
                else:
                    rc = _raw(*args)
                if rc == 0:
                    LAUNCHES[0] += 1
                return rc
            self._cache[name] = fn
        return fn

# every symbol include/idrk.h declares (tests check the library exports all of them)
EXPORTS = ["idrk_version", "idrk_device_sm_count", "idrk_hash_encode_fwd", "idrk_hash_encode_bwd",
           "idrk_posenc_fwd", "idrk_posenc_bwd", "idrk_gemm", "idrk_split_tf32", "idrk_weight_norm_fwd",
           "idrk_weight_norm_bwd", "idrk_colsum", "idrk_sdf_head", "idrk_sdf_squash",
           "idrk_rt_init", "idrk_rt_top", "idrk_rt_step", "idrk_rt_linesearch", "idrk_rt_end",
           "idrk_rt_select_sampler", "idrk_rt_sampler_points", "idrk_rt_sampler_resolve", "idrk_rt_secant",
           "idrk_rt_select_minsdf", "idrk_rt_minsdf_points", "idrk_rt_minsdf_resolve", "idrk_rt_chunk_counts",
           "idrk_sumsq", "idrk_clip_adam", "idrk_act_bwd", "idrk_gemm_f16s", "idrk_split_f16", "idrk_nffb_encode_fwd",
           "idrk_hash_encode_f16pair", "idrk_camera_rays", "idrk_idr_loss", "idrk_scale3",
           "idrk_fourier_dx_fwd", "idrk_fourier_dx_bwd", "idrk_morton_sort_workspace", "idrk_morton_sort",
           "idrk_hash_encode_bwd_det_workspace", "idrk_hash_encode_bwd_det", "idrk_sdf_squash_rows", "idrk_sumsq_det",
           "idrk_rt_linesearch_points", "idrk_rt_linesearch_resolve",
           "idrk_gemm_p16", "idrk_split_p16", "idrk_weight_norm_fwd_p16", "idrk_act_bwd_p16", "idrk_nffb_encode_f16pair", "idrk_rt_iter_tail", "idrk_posenc_dx_bwd"]


class NffbDesc(ctypes.Structure):
    """Mirror of idrk_nffb_t."""
    _fields_ = [("grid", HashGridDesc), ("bands", ctypes.c_float * 32),
                ("n_bands", ctypes.c_int32), ("include_input", ctypes.c_int32), ("n_lin", ctypes.c_int32),
                ("width", ctypes.c_int32), ("chunk", ctypes.c_int32), ("style", ctypes.c_int32),
                ("n_levels_div", ctypes.c_int32), ("bound", ctypes.c_float), ("w0", ctypes.c_float), ("eps", ctypes.c_float),
                ("lin_w", ctypes.c_void_p * 16), ("lin_b", ctypes.c_void_p * 16),
                ("out_w", ctypes.c_void_p), ("out_b", ctypes.c_void_p), ("style_w", ctypes.c_void_p), ("style_b", ctypes.c_void_p)]


class EpilogueH(ctypes.Structure):
    """Mirror of idrk_epilogue_f16_t."""
    _fields_ = [("C", ctypes.c_void_p), ("C_h", ctypes.c_void_p), ("C_l", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("ldc", ctypes.c_int32), ("ldh", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("act_param", ctypes.c_float), ("scale", ctypes.c_float),
                ("dot_w", ctypes.c_void_p), ("dot_out", ctypes.c_void_p), ("ld_dot", ctypes.c_int32)]


class EpilogueP(ctypes.Structure):
    """Mirror of idrk_epilogue_p16_t."""
    _fields_ = [("C", ctypes.c_void_p), ("S", ctypes.c_void_p), ("C_h", ctypes.c_void_p), ("C_l", ctypes.c_void_p),
                ("bias", ctypes.c_void_p), ("aux", ctypes.c_void_p),
                ("ldc", ctypes.c_int32), ("lds", ctypes.c_int32), ("ldh", ctypes.c_int32), ("ldaux", ctypes.c_int32),
                ("c_fmt", ctypes.c_int32), ("mode", ctypes.c_int32), ("act_param", ctypes.c_float), ("scale", ctypes.c_float),
                ("accumulate", ctypes.c_int32)]


P16_FP16, P16_BF16 = 0, 1


class RayStateDesc(ctypes.Structure):
    """Mirror of idrk_ray_state_t."""
    _fields_ = [("cam_loc", ctypes.c_void_p), ("ray_dirs", ctypes.c_void_p),
                ("n_rays", ctypes.c_int32), ("num_pixels", ctypes.c_int32)] + \
               [(n, ctypes.c_void_p) for n in ("t0", "t1", "cur_s", "cur_e", "nxt_s", "nxt_e", "ps", "pe", "min_dis",
                                               "max_dis", "unf_s", "unf_e", "slot_s", "slot_e")]

GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
PREC_FP32, PREC_TF32, PREC_3XTF32 = 0, 1, 3
EPI_NONE, EPI_SOFTPLUS, EPI_RELU, EPI_MUL_AUX, EPI_SINE, EPI_TANH = 0, 1, 2, 3, 4, 5


class Epilogue(ctypes.Structure):
    """Mirror of idrk_epilogue_t."""
    _fields_ = [
        ("C", ctypes.c_void_p), ("C_hi", ctypes.c_void_p), ("C_lo", ctypes.c_void_p), ("S", ctypes.c_void_p),
        ("bias", ctypes.c_void_p), ("aux", ctypes.c_void_p),
        ("ldc", ctypes.c_int32), ("lds", ctypes.c_int32), ("ldaux", ctypes.c_int32),
        ("mode", ctypes.c_int32), ("act_param", ctypes.c_float), ("scale", ctypes.c_float),
        ("accumulate", ctypes.c_int32),
    ]

_ARG_ERRORS = {-1: "bad argument", -2: "pointer or leading dimension not 16-byte aligned",
               -3: "unsupported configuration", -4: "CUDA driver entry point unavailable"}


def lib():
    """Loads libidrk.so on first use; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IdrkError("libidrk.so is missing (%s). Build it with `python hashmodnffbanks-idr_b200/csrc/build.py`; "
                            "there is no CPU or PyTorch fallback." % LIB_PATH)
        cdll = ctypes.CDLL(LIB_PATH)
        _declare(cdll)
        _lib = _Proxy(cdll)
    return _lib


def _declare(L):
    c = ctypes
    vp, i32, i64, f32 = c.c_void_p, c.c_int32, c.c_int64, c.c_float
    L.idrk_version.restype = c.c_int
    L.idrk_device_sm_count.argtypes = [c.POINTER(c.c_int)]
    L.idrk_hash_encode_fwd.argtypes = [c.POINTER(HashGridDesc), vp, i64, i32, vp, i32, vp, vp, vp, vp]
    L.idrk_hash_encode_bwd.argtypes = [c.POINTER(HashGridDesc), vp, i64, i32, vp, i32, c.POINTER(vp), vp, i32, vp, vp]
    fp = c.POINTER(c.c_float)
    L.idrk_posenc_fwd.argtypes = [vp, i64, i32, i32, fp, i32, i32, vp, i32, vp]
    L.idrk_posenc_bwd.argtypes = [vp, i64, i32, i32, fp, i32, i32, vp, i32, vp, i32, vp]
    L.idrk_posenc_dx_bwd.argtypes = [vp, i32, vp, i64, i32, i32, fp, i32, i32, vp, i32, vp, i32, vp, i32, vp]
    L.idrk_gemm.argtypes = [i32, i32, i64, i32, i32, vp, vp, i32, vp, vp, i32, c.POINTER(Epilogue), vp, i32, vp]
    L.idrk_split_tf32.argtypes = [vp, i64, i32, i32, f32, vp, vp, i32, i32, vp, vp]
    L.idrk_weight_norm_fwd.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp]
    L.idrk_weight_norm_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, i32, vp]
    L.idrk_colsum.argtypes = [vp, i64, i32, i32, vp, vp]
    L.idrk_sdf_head.argtypes = [vp, i64, i32, i32, vp, vp, f32, vp, vp, vp]
    L.idrk_sdf_squash.argtypes = [vp, i64, f32, vp, vp, vp]
    L.idrk_sdf_squash_rows.argtypes = [vp, i64, i32, i32, f32, vp, i32, vp, vp, vp]
    rs = c.POINTER(RayStateDesc)
    L.idrk_rt_init.argtypes = [rs, vp, vp, vp, vp, vp]
    L.idrk_rt_top.argtypes = [rs, vp, i32, f32, vp, vp]
    L.idrk_rt_step.argtypes = [rs, vp, vp, vp, vp]
    L.idrk_rt_linesearch.argtypes = [rs, vp, vp, i32, f32, vp, vp, vp]
    L.idrk_rt_end.argtypes = [rs, vp, vp, i32, vp]
    L.idrk_rt_linesearch_points.argtypes = [rs, vp, vp, fp, i32, vp, vp, vp]
    L.idrk_rt_linesearch_resolve.argtypes = [rs, vp, vp, fp, i32, vp]
    L.idrk_rt_iter_tail.argtypes = [rs, vp, vp, fp, i32, f32, vp, vp]
    L.idrk_rt_select_sampler.argtypes = [rs, vp, vp, vp, vp]
    L.idrk_rt_sampler_points.argtypes = [rs, vp, i32, i32, i32, vp, vp, vp, vp]
    L.idrk_rt_chunk_counts.argtypes = [vp, i32, i32, i32, vp, vp]
    L.idrk_rt_sampler_resolve.argtypes = [rs, vp, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.idrk_rt_secant.argtypes = [rs, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.idrk_rt_select_minsdf.argtypes = [rs, vp, vp, vp, vp, vp, vp, vp]
    L.idrk_rt_minsdf_points.argtypes = [rs, vp, i32, i32, i32, vp, vp, vp, vp]
    L.idrk_rt_minsdf_resolve.argtypes = [rs, vp, i32, i32, vp, vp, vp, vp]
    L.idrk_act_bwd.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, i64, i32, i32, f32, f32, vp, vp, vp, i32, vp]
    L.idrk_gemm_p16.argtypes = [i32, i64, i32, i32, vp, vp, i32, i32, vp, vp, i32, i32, c.POINTER(EpilogueP), vp, i32, vp]
    L.idrk_split_p16.argtypes = [vp, i64, i32, i32, f32, vp, vp, i32, i32, i32, vp, vp]
    L.idrk_weight_norm_fwd_p16.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp, vp, i32, i32, vp]
    L.idrk_act_bwd_p16.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, i64, i32, i32, f32, f32, vp, i32, vp, vp, i32, i32, vp]
    L.idrk_gemm_f16s.argtypes = [i64, i32, i32, vp, vp, i32, vp, vp, i32, c.POINTER(EpilogueH), vp, vp]
    L.idrk_split_f16.argtypes = [vp, i64, i32, i32, f32, vp, vp, i32, i32, vp, vp, i32, i32, f32, vp, vp]
    L.idrk_nffb_encode_fwd.argtypes = [c.POINTER(NffbDesc), vp, i64, i32, vp, i32, vp, vp]
    L.idrk_nffb_encode_f16pair.argtypes = [c.POINTER(NffbDesc), vp, i64, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, f32, vp]
    L.idrk_hash_encode_f16pair.argtypes = [c.POINTER(HashGridDesc), vp, i64, i32, vp, vp, vp, i32, i32, vp, vp, i32, i32, f32, vp]
    L.idrk_camera_rays.argtypes = [vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp]
    L.idrk_idr_loss.argtypes = [vp, i32, vp, vp, vp, vp, i32, i64, vp, i32, i64, f32, f32, f32, vp, vp, vp, vp, vp]
    L.idrk_fourier_dx_fwd.argtypes = [vp, i32, vp, i32, vp, i32, i64, vp, vp]
    L.idrk_fourier_dx_bwd.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, i64, vp, i32, i32, vp, vp]
    L.idrk_hash_encode_bwd_det_workspace.argtypes = [c.POINTER(HashGridDesc), i64, c.POINTER(c.c_int64)]
    L.idrk_hash_encode_bwd_det.argtypes = [c.POINTER(HashGridDesc), vp, i64, i32, vp, i32, c.POINTER(vp), vp, i64, vp]
    L.idrk_morton_sort_workspace.argtypes = [i64, c.POINTER(c.c_int64)]
    L.idrk_morton_sort.argtypes = [vp, i64, i32, fp, fp, i32, vp, vp, i64, vp]
    L.idrk_scale3.argtypes = [vp, vp, vp, i64, vp, vp, i64, vp, vp, i64, vp]
    L.idrk_sumsq.argtypes = [vp, i64, vp, vp]
    L.idrk_sumsq_det.argtypes = [vp, i64, vp, vp, i32, vp]
    L.idrk_clip_adam.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp, f32, vp]
    for fn in EXPORTS:
        getattr(L, fn).restype = c.c_int


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise IdrkError("%s: %s (code %d)" % (what, _ARG_ERRORS.get(rc, "error"), rc))
    raise IdrkError("%s: CUDA error %d" % (what, rc))


def require_cuda(t: torch.Tensor, name="tensor"):
    if not t.is_cuda:
        raise IdrkError("%s must live on a CUDA device: idrk has no CPU path" % name)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
