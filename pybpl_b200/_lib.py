"""ctypes binding of libpybpl_b200.so (include/pybpl_b200.h).

The library is the product: if it is missing or fails to load, importing anything that computes
raises -- there is no CPU or PyTorch fallback for the render+score path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpybpl_b200.so")

c_int_p = ctypes.POINTER(ctypes.c_int)
c_float_p = ctypes.POINTER(ctypes.c_float)
vp = ctypes.c_void_p

IMSIZE = 105
IMG_WORDS = 345
NCPT = 5
MAX_SUBSTROKES = 128
MAX_STROKES = 32
MAX_TABLE = 1024


class Params(ctypes.Structure):
    _fields_ = [("imsize", ctypes.c_int), ("fsize", ctypes.c_int), ("ink_ncon", ctypes.c_int),
                ("ink_pp", ctypes.c_float), ("ink_max_dist", ctypes.c_float), ("ink_a", ctypes.c_float),
                ("ink_b", ctypes.c_float), ("hinton", ctypes.c_int)]


class TokenBatchC(ctypes.Structure):
    _fields_ = [("n_tokens", ctypes.c_int), ("ncpt", ctypes.c_int), ("neval", ctypes.c_int),
                ("tok_stroke_off", vp), ("stroke_sub_off", vp), ("ctrl", vp), ("invscale", vp), ("position", vp),
                ("affine", vp), ("eps", vp), ("sigma", vp), ("basis", vp), ("order", vp), ("cycles_out", vp)]


class StrokeBatchC(ctypes.Structure):
    _fields_ = [("n_tokens", ctypes.c_int), ("tok_sub_off", vp), ("sub_pt_off", vp), ("pts", vp), ("eps", vp),
                ("sigma", vp)]


class Targets(ctypes.Structure):
    _fields_ = [("image_bits", vp), ("img_idx", vp), ("n_images", ctypes.c_int), ("all_pairs", ctypes.c_int)]


class FwdOut(ctypes.Structure):
    _fields_ = [("ll", vp), ("off_page", vp), ("pimg", vp)]


class AnyGeomC(ctypes.Structure):       # bpl_any_geom: rendering for any Parameters.imsize
    _fields_ = [("H", ctypes.c_int), ("W", ctypes.c_int), ("images", vp), ("img_idx", vp), ("workspace", vp),
                ("workspace_floats", ctypes.c_longlong)]


class Grads(ctypes.Structure):
    _fields_ = [("g_ll", vp), ("g_pimg", vp), ("g_ctrl", vp), ("g_invscale", vp), ("g_position", vp),
                ("g_affine", vp), ("g_pts", vp), ("g_eps", vp), ("g_sigma", vp)]


ADAM_MAX_TENSORS = 8


class AdamArgs(ctypes.Structure):
    _fields_ = [("n_tensors", ctypes.c_int), ("maximize", ctypes.c_int), ("lr", ctypes.c_float), ("beta1", ctypes.c_float),
                ("beta2", ctypes.c_float), ("eps", ctypes.c_float), ("state", vp), ("param", vp * ADAM_MAX_TENSORS),
                ("grad", vp * ADAM_MAX_TENSORS), ("exp_avg", vp * ADAM_MAX_TENSORS), ("exp_avg_sq", vp * ADAM_MAX_TENSORS),
                ("n", ctypes.c_longlong * ADAM_MAX_TENSORS), ("lo", ctypes.c_float * ADAM_MAX_TENSORS),
                ("hi", ctypes.c_float * ADAM_MAX_TENSORS)]


class TokenPriorParams(ctypes.Structure):
    _fields_ = [("sigma_shape", ctypes.c_float), ("sigma_invscale", ctypes.c_float), ("sigma_attach", ctypes.c_float),
                ("sigma_x", ctypes.c_float), ("sigma_y", ctypes.c_float), ("min_blur_sigma", ctypes.c_float),
                ("max_blur_sigma", ctypes.c_float)]


class TokenTypeC(ctypes.Structure):
    _fields_ = [("ctrl_type", vp), ("invscale_type", vp), ("rel_cat", vp), ("attach_ix", vp), ("attach_subix", vp),
                ("gpos", vp), ("eval_spot_type", vp), ("eval_spot", vp)]


class TokenPriorGrads(ctypes.Structure):
    _fields_ = [("g_ctrl", vp), ("g_invscale", vp), ("g_position", vp), ("g_eval_spot", vp), ("g_ctrl_type", vp),
                ("g_invscale_type", vp), ("g_gpos", vp), ("g_eval_spot_type", vp)]


class TypeLibC(ctypes.Structure):
    _fields_ = [("n_prim", ctypes.c_int), ("kmax", ctypes.c_int), ("nsub_max", ctypes.c_int), ("n_hist", ctypes.c_int),
                ("hist_bins", ctypes.c_int), ("pkappa", vp), ("pmat_nsub", vp), ("cdf_start", vp), ("cdf_T", vp),
                ("shape_mu", vp), ("shape_L", vp), ("gamma_con", vp), ("gamma_rate", vp), ("rel_mixprob", vp), ("sp_cdf", vp),
                ("sp_xlab", vp), ("sp_ylab", vp)]


class TypeLogTablesC(ctypes.Structure):
    _fields_ = [("logp_kappa", vp), ("logp_nsub", vp), ("logp_start", vp), ("logp_T", vp), ("logp_mix", vp), ("sp_logp", vp),
                ("sp_log_binarea", ctypes.c_float)]


class TypeOutC(ctypes.Structure):
    _fields_ = [("stroke_nsub", vp), ("subid", vp), ("ctrl_type", vp), ("invscale_type", vp), ("rel_cat", vp),
                ("attach_ix", vp), ("attach_subix", vp), ("gpos", vp), ("eval_spot_type", vp)]


EXPORTS = {
    "bpl_version": (ctypes.c_int, []),
    "bpl_last_error": (ctypes.c_char_p, []),
    "bpl_default_params": (None, [ctypes.POINTER(Params)]),
    "bpl_device_info": (ctypes.c_int, [c_int_p, c_int_p]),
    "bpl_basis_table_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, vp]),
    "bpl_basis_rows_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, vp, vp]),
    "bpl_spline_eval": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "bpl_spline_eval_bwd": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "bpl_spline_adaptive_neval": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_float,
                                                 ctypes.c_int, vp, vp, vp]),
    "bpl_lstsq_operator_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, vp, vp]),
    "bpl_spline_residual": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]),
    "bpl_vanilla_to_motor": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "bpl_pack_images_u8": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp]),
    "bpl_pack_images_f32": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp]),
    "bpl_score_tokens": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(TokenBatchC), ctypes.POINTER(Targets),
                                        ctypes.POINTER(FwdOut), vp]),
    "bpl_score_tokens_lse": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(TokenBatchC), ctypes.POINTER(Targets),
                                            ctypes.POINTER(FwdOut), vp, vp, ctypes.c_int, vp]),
    "bpl_order_tokens": (ctypes.c_int, [ctypes.POINTER(TokenBatchC), vp, vp]),
    "bpl_order_tokens_by_cycles": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp]),
    "bpl_adam_step": (ctypes.c_int, [ctypes.POINTER(AdamArgs), vp]),
    "bpl_scale_token_grads": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.POINTER(Grads),
                                             ctypes.POINTER(Grads), vp]),
    "bpl_order_tokens_by_cost": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(TokenBatchC), vp, vp, vp]),
    "bpl_score_tokens_grad": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(TokenBatchC),
                                             ctypes.POINTER(Targets), ctypes.POINTER(FwdOut), ctypes.POINTER(Grads), vp]),
    "bpl_render_strokes": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(StrokeBatchC),
                                          ctypes.POINTER(Targets), ctypes.POINTER(FwdOut), vp]),
    "bpl_render_strokes_grad": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(StrokeBatchC),
                                               ctypes.POINTER(Targets), ctypes.POINTER(FwdOut), ctypes.POINTER(Grads), vp]),
    "bpl_render_any_workspace": (ctypes.c_longlong, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "bpl_render_strokes_any": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(AnyGeomC), ctypes.POINTER(StrokeBatchC),
                                              ctypes.POINTER(FwdOut), vp]),
    "bpl_render_strokes_any_grad": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(AnyGeomC),
                                                   ctypes.POINTER(StrokeBatchC), ctypes.POINTER(FwdOut),
                                                   ctypes.POINTER(Grads), vp]),
    "bpl_token_points": (ctypes.c_int, [ctypes.POINTER(Params), ctypes.POINTER(TokenBatchC), vp, vp, vp, vp, vp, vp]),
    "bpl_score_token_prior": (ctypes.c_int, [ctypes.POINTER(TokenPriorParams), ctypes.POINTER(TokenBatchC),
                                             ctypes.POINTER(TokenTypeC), vp, ctypes.POINTER(TokenPriorGrads), vp]),
    "bpl_sample_tokens": (ctypes.c_int, [ctypes.POINTER(TokenPriorParams), ctypes.POINTER(TokenBatchC),
                                         ctypes.POINTER(TokenTypeC), ctypes.c_ulonglong, ctypes.c_ulonglong, vp, vp, vp, vp, vp]),
    "bpl_sample_types": (ctypes.c_int, [ctypes.POINTER(TypeLibC), ctypes.c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, vp, vp,
                                        vp, vp, ctypes.POINTER(TypeOutC), vp]),
    "bpl_score_type_prior": (ctypes.c_int, [ctypes.POINTER(TypeLibC), ctypes.POINTER(TypeLogTablesC), ctypes.c_int, vp, vp, vp,
                                            ctypes.POINTER(TokenTypeC), vp, vp, vp, vp]),
    "bpl_logsumexp_partial": (ctypes.c_int, [vp, ctypes.c_int, vp, vp]),
    "bpl_logsumexp_combine": (ctypes.c_int, [vp, ctypes.c_int, vp, vp]),
    "bpl_fp32_probe": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp]),
    "bpl_smem_probe": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp]),
    "bpl_launch_count": (ctypes.c_longlong, []),
}

_lib = None


LSE_SCRATCH_CTAS = 2048      # BPL_LSE_SCRATCH_CTAS (include/pybpl_b200.h)


class BplError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (python -m pybpl_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise BplError(f"{SO_PATH} is missing: build it with `python -m pybpl_b200.build` "
                           "(there is no CPU fallback for the render+score path)")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in EXPORTS.items():
            f = getattr(L, name)       # AttributeError if the .so does not export what the header declares
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != 0:
        msg = lib().bpl_last_error()
        raise BplError(f"libpybpl_b200 error {code}: {msg.decode() if msg else ''}")


def default_params():
    ps = Params()
    lib().bpl_default_params(ctypes.byref(ps))
    return ps


def launch_count():
    return int(lib().bpl_launch_count())
