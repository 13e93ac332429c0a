"""ctypes binding of the C-ABI library ``libvxb200.so`` (declared in include/vxb200.h).

There is no fallback: if the library is missing or the device is not sm_100, loading raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvxb200.so")

_lib = None

c_int, c_ll, c_float, c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p


class VxError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VxError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU/PyTorch fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vx_last_error.restype = ctypes.c_char_p
    return _lib


LAUNCHES = 0        # kernels launched through the C ABI by this process (one per successful op call)
_TIMING = None      # when set to a list, every op call is bracketed by CUDA events: (name, start, end, meta)


def check(rc, what=""):
    global LAUNCHES
    if rc != 0:
        raise VxError(f"{what}: {lib().vx_last_error().decode()}")
    LAUNCHES += 1


def ptr(t):
    """Raw device/host pointer of a torch tensor (or None)."""
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_sm100():
    check(lib().vx_require_sm100(), "vx_require_sm100")


_WARNED = set()


def note_compute_dtype(model_dtype, what):
    """The kernels compute in bf16 with fp32 accumulation, whatever dtype the caller's modules hold.  The reference CLI
    defaults to fp16 (inference.py:44,150-151): such a model is accepted -- weights are re-rounded to bf16 (same 16 bits,
    3 fewer mantissa bits, no overflow risk), inputs / outputs are converted at the boundary -- but never silently."""
    import warnings
    import torch
    if model_dtype != torch.bfloat16 and (what, model_dtype) not in _WARNED:
        _WARNED.add((what, model_dtype))
        warnings.warn(f"vexpress_b200: {what} holds {model_dtype} parameters; the sm_100a kernels compute in bfloat16 "
                      f"(fp32 accumulation): weights are re-rounded to bf16, inputs/outputs converted at the boundary. "
                      f"Use .to(torch.bfloat16) to make this explicit.", stacklevel=3)
