"""ctypes binding of libipdm_b200.so (the C ABI declared in include/ipdm_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`inverseproblemwithdiffusionmodel_b200/csrc/build.sh`.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_ulonglong, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
ABI_VERSION = 3     # what this binding was written against (include/ipdm_b200.h); lib() refuses another library
LIB_PATH = os.environ.get("IPDM_B200_LIB") or os.path.join(_HERE, "libipdm_b200.so")   # override: A/B builds of the same ABI


class AldScalars(ctypes.Structure):
    """mirror of `ipdm_ald_scalars`"""
    _fields_ = [("step", c_float), ("noise_scale", c_float), ("kappa", c_float), ("sigma", c_float)]


class ConvDesc(ctypes.Structure):
    """mirror of `ipdm_conv_desc`"""
    _fields_ = [("in_f16", c_void_p), ("w_f16", c_void_p), ("bias", c_void_p), ("residual", c_void_p),
                ("out_f32", c_void_p), ("out_f16", c_void_p), ("stats", c_void_p),
                ("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
                ("taps", c_int), ("dilation", c_int), ("flags", c_int), ("slices", c_int), ("slice_shift", c_int),
                ("residual_f16", c_void_p), ("out_raw_f16", c_void_p), ("acc_scale", ctypes.c_float), ("out_f16_scale", ctypes.c_float), ("tap_mask", ctypes.c_uint16 * 16)]


class Rng(ctypes.Structure):
    """mirror of `ipdm_rng`"""
    _fields_ = [("seed", c_uint64), ("seed_dev", c_void_p), ("rng_step", c_uint32), ("chain_base", ctypes.c_int32),
                ("chain_ids", c_void_p), ("chain_elems", c_size_t)]


def rng(seed=0, rng_step=0, chain_ids=None, chain_elems=0, seed_dev=None, chain_base=0):
    """`ipdm_rng` for one call.  chain_ids: int32 device tensor of GLOBAL chain ids (one per sample) or None
    (sample i is chain chain_base + i); seed_dev: uint64/int64 device tensor[1] XORed into `seed` on the device."""
    return Rng(int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(seed_dev), int(rng_step) & 0xFFFFFFFF, int(chain_base), ptr(chain_ids),
               int(chain_elems))


CONV_F16_ELU, CONV_F16_PRE_RES, CONV_RES_ELU, CONV_POOL2 = 1, 2, 4, 8

# name -> (restype, argtypes); every symbol of include/ipdm_b200.h
SIGNATURES = {
    "ipdm_abi_version": (c_int, []),
    "ipdm_last_error": (c_char_p, []),
    "ipdm_launch_count": (c_ulonglong, []),
    "ipdm_sense_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ipdm_sense_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ipdm_sense_adjoint": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ipdm_kspace_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_caxpy": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_size_t, c_void_p]),
    "ipdm_sense_plan_create": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]),
    "ipdm_sense_plan_destroy": (c_int, [c_void_p]),
    "ipdm_sense_plan_info": (c_int, [c_void_p, POINTER(c_int)]),
    "ipdm_sense_forward_plan": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ipdm_sense_adjoint_plan": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ipdm_langevin_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(AldScalars), c_void_p, c_void_p,
                                     c_void_p, c_size_t, POINTER(Rng), c_void_p]),
    "ipdm_ald_sense_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_int, POINTER(AldScalars), c_void_p, c_void_p, POINTER(Rng), c_void_p]),
    "ipdm_ald_sense_step_plan": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         POINTER(AldScalars), c_void_p, c_void_p, POINTER(Rng), c_void_p]),
    "ipdm_ald_advance": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ipdm_temporal_tv_step": (c_int, [c_void_p, c_int, c_int, c_size_t, c_float, c_void_p]),
    "ipdm_planar_to_c64": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "ipdm_c64_to_planar": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "ipdm_chain_stats_accumulate": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p]),
    "ipdm_image_sums": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p]),
    "ipdm_ssim": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p]),
    "ipdm_conv_igemm": (c_int, [POINTER(ConvDesc), c_void_p]),
    "ipdm_conv_direct": (c_int, [POINTER(ConvDesc), c_void_p]),
    "ipdm_debug_option": (c_int, [c_int, c_int]),
    "ipdm_conv_first": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_conv_last": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_instnorm_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_instnorm_apply_elu": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ipdm_conv_first_f16out": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_instnorm_apply_elu_f16in": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ipdm_bilinear_add_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_f16_range_audit": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ipdm_instnorm_apply_elu_s2d": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_act_to_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "ipdm_act_to_f16_scaled": (c_int, [c_void_p, c_void_p, c_size_t, c_int, ctypes.c_float, c_void_p]),
    "ipdm_maxpool5_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_bilinear_add": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_meanpool2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_meanpool2_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_pack_weights_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ipdm_maxpool5_slices_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_void_p]),
    "ipdm_conv3d_first": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_conv3d_last": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_gather_t_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_interleave_t": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_void_p]),
    "ipdm_add_act": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "ipdm_patch_fold": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ipdm_patch_fold_sched": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
}

_lib = None


class IpdmError(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IpdmError(f"{LIB_PATH} not found: build the CUDA extension first (there is no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        got = handle.ipdm_abi_version()
        if got != ABI_VERSION:
            raise IpdmError(f"{LIB_PATH} has ABI version {got}, this binding needs {ABI_VERSION}: rebuild the library (csrc/build.sh)")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ipdm_last_error().decode()
        raise IpdmError(f"{what or 'ipdm call'} failed (rc={rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def fresh_seed():
    """A seed drawn from torch's global generator -- what a sampler uses when the caller passes none, so that
    successive calls draw fresh noise (like the reference's torch.randn_like) and torch.manual_seed reproduces a run."""
    import torch
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class SensePlan:
    """Owner of one `ipdm_sense_plan_*` handle (a column mask compiled for one device and one (H, W))."""

    def __init__(self, mask_u8_host, H, W):
        import numpy as np
        m = np.ascontiguousarray(mask_u8_host, dtype=np.uint8).reshape(-1, W)
        self.frames, self.H, self.W = int(m.shape[0]), int(H), int(W)
        handle = c_void_p()
        check(lib().ipdm_sense_plan_create(m.ctypes.data, self.frames, self.H, self.W, ctypes.byref(handle)), "sense_plan_create")
        self.handle = handle
        info = (c_int * 8)()
        check(lib().ipdm_sense_plan_info(self.handle, info), "sense_plan_info")
        self.pruned, self.ns_max, self.ns_pad, self.groups_max = bool(info[0]), info[1], info[2], info[3]
        self.pruned_rows = bool(info[7])

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.ipdm_sense_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise IpdmError("ipdm_b200 runs on CUDA tensors only (no CPU fallback): got a tensor on " + str(t.device))
