"""ctypes binding of libb200q.so (include/b200q.h).  There is no fallback: if the shared library
is missing the import fails loudly -- the product path is the CUDA engine or nothing."""
import ctypes
import os

from ._build import LIB

B200Q_OK = 0
LAYOUT_GPTQ, LAYOUT_AWQ_GEMM, LAYOUT_MARLIN, LAYOUT_HQQ, LAYOUT_AWQ_GEMV, LAYOUT_ORT = 0, 1, 2, 3, 4, 5
KERNEL_GEMV, KERNEL_GEMM, KERNEL_GENERIC = 1, 2, 3
PEER_Y_TAGGED, PEER_X_TAGGED = 1, 2
PEER_NODE_EPOCH = 4


class Layer(ctypes.Structure):
    """struct b200q_layer"""
    _fields_ = [("layout", ctypes.c_int32), ("bits", ctypes.c_int32), ("group_size", ctypes.c_int32),
                ("K", ctypes.c_int32), ("N", ctypes.c_int32), ("zero_bias", ctypes.c_int32),
                ("qweight", ctypes.c_void_p), ("qzeros", ctypes.c_void_p), ("scales", ctypes.c_void_p),
                ("g_idx", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("x_perm", ctypes.c_void_p)]


class ChainStep(ctypes.Structure):
    """struct b200q_chain_step"""
    _fields_ = [("layers", ctypes.POINTER(ctypes.POINTER(Layer))), ("n_layers", ctypes.c_int32), ("x", ctypes.c_void_p),
                ("ldx", ctypes.c_int64), ("y", ctypes.POINTER(ctypes.c_void_p)), ("ldy", ctypes.POINTER(ctypes.c_int64))]


class Fusion(ctypes.Structure):
    """struct b200q_fusion"""
    _fields_ = [("x_mul", ctypes.c_void_p), ("residual", ctypes.c_void_p), ("ldres", ctypes.c_int64),
                ("act_dtype", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class PeerSync(ctypes.Structure):
    """struct b200q_peer_sync"""
    _fields_ = [("n_peers", ctypes.c_int32), ("self_rank", ctypes.c_int32), ("counters", ctypes.POINTER(ctypes.c_void_p)),
                ("epoch", ctypes.c_void_p), ("wait_slot", ctypes.c_int32), ("wait_count", ctypes.c_uint32),
                ("post_slot", ctypes.c_int32), ("flags", ctypes.c_uint32),
                ("tag_stride", ctypes.c_uint32), ("y_seq", ctypes.c_uint32), ("x_seq", ctypes.c_uint32)]


EXPORTS = ["b200q_linear", "b200q_linear_group", "b200q_gemv", "b200q_gemm", "b200q_linear_sharded", "b200q_dequant", "b200q_unpack",
           "b200q_workspace_bytes", "b200q_gemv_max_m", "b200q_select_kernel", "b200q_launch_count",
           "b200q_strerror", "b200q_last_cuda_error", "b200q_version", "b200q_debug_set_timeline", "b200q_repack_gptq4", "b200q_debug_decode_plan", "b200q_debug_set_option", "b200q_linear_group_sharded", "b200q_sharded_posts",
           "b200q_peer_epoch_advance", "b200q_peer_wait", "b200q_peer_untag", "b200q_repack_actorder",
           "b200q_repack_from_gptq4", "b200q_linear_ex", "b200q_workspace_bytes_ex", "b200q_chain_plan_bytes", "b200q_chain_plan", "b200q_chain_run", "b200q_debug_set_chain_timeline"]


def _load():
    if not os.path.exists(LIB):
        raise ImportError(
            f"{LIB} not found: build it with `python -m qllm_b200._build` (nvcc, sm_100a). "
            "qllm_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB)
    if os.environ.get("B200Q_LIB"):                       # A/B run against an older build: stub what it lacks
        class _Missing:
            argtypes = restype = None

            def __call__(self, *a):
                raise RuntimeError("entry point missing from the B200Q_LIB build")
        for name in EXPORTS:
            if not hasattr(lib, name):
                setattr(lib, name, _Missing())
    P, I64, SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
    LP = ctypes.POINTER(Layer)
    fwd = [LP, P, I64, I64, P, I64, P, SZ, P]
    for name in ("b200q_linear", "b200q_gemv", "b200q_gemm"):
        getattr(lib, name).argtypes = fwd
        getattr(lib, name).restype = ctypes.c_int
    lib.b200q_linear_ex.argtypes = [LP, P, I64, I64, P, I64, ctypes.POINTER(Fusion), P, SZ, P]
    lib.b200q_linear_ex.restype = ctypes.c_int
    lib.b200q_workspace_bytes_ex.argtypes = [LP, I64, ctypes.POINTER(Fusion)]
    lib.b200q_workspace_bytes_ex.restype = SZ
    lib.b200q_linear_sharded.argtypes = [LP, P, I64, I64, ctypes.POINTER(P), ctypes.c_int32, I64, I64, P, SZ, P]
    lib.b200q_linear_sharded.restype = ctypes.c_int
    lib.b200q_linear_group.argtypes = [ctypes.POINTER(LP), ctypes.c_int32, P, I64, I64, ctypes.POINTER(P),
                                       ctypes.POINTER(I64), P, SZ, P]
    lib.b200q_linear_group.restype = ctypes.c_int
    lib.b200q_linear_group_sharded.argtypes = [ctypes.POINTER(LP), ctypes.c_int32, P, I64, I64, ctypes.POINTER(P),
                                               ctypes.POINTER(I64), ctypes.POINTER(I64), ctypes.POINTER(PeerSync), P, SZ, P]
    lib.b200q_linear_group_sharded.restype = ctypes.c_int
    lib.b200q_sharded_posts.argtypes = [ctypes.POINTER(LP), ctypes.c_int32, I64]
    lib.b200q_sharded_posts.restype = ctypes.c_int
    lib.b200q_peer_untag.argtypes = [P, I64, P, I64, I64, I64, ctypes.POINTER(PeerSync), P]
    lib.b200q_peer_untag.restype = ctypes.c_int
    lib.b200q_peer_wait.argtypes = [ctypes.POINTER(PeerSync), P]
    lib.b200q_peer_wait.restype = ctypes.c_int
    lib.b200q_peer_epoch_advance.argtypes = [P, P]
    lib.b200q_peer_epoch_advance.restype = ctypes.c_int
    lib.b200q_chain_plan_bytes.argtypes = [ctypes.POINTER(ChainStep), ctypes.c_int32, I64]
    lib.b200q_chain_plan_bytes.restype = SZ
    lib.b200q_chain_plan.argtypes = [ctypes.POINTER(ChainStep), ctypes.c_int32, I64, P, SZ, ctypes.POINTER(SZ)]
    lib.b200q_chain_plan.restype = ctypes.c_int
    lib.b200q_chain_run.argtypes = [P, P, P, SZ, P]
    lib.b200q_chain_run.restype = ctypes.c_int
    lib.b200q_dequant.argtypes = [LP, P, P]
    lib.b200q_dequant.restype = ctypes.c_int
    lib.b200q_unpack.argtypes = [LP, P, P, P]
    lib.b200q_unpack.restype = ctypes.c_int
    lib.b200q_repack_actorder.argtypes = [LP, P, P, P]
    lib.b200q_repack_actorder.restype = ctypes.c_int
    lib.b200q_repack_gptq4.argtypes = [LP, P, P, P, P]
    lib.b200q_repack_gptq4.restype = ctypes.c_int
    lib.b200q_repack_from_gptq4.argtypes = [LP, ctypes.c_int32, P, P, P, P]
    lib.b200q_repack_from_gptq4.restype = ctypes.c_int
    lib.b200q_workspace_bytes.argtypes = [LP, I64]
    lib.b200q_workspace_bytes.restype = SZ
    lib.b200q_gemv_max_m.restype = ctypes.c_int
    lib.b200q_select_kernel.argtypes = [LP, I64]
    lib.b200q_select_kernel.restype = ctypes.c_int
    lib.b200q_launch_count.restype = ctypes.c_uint64
    lib.b200q_strerror.argtypes = [ctypes.c_int]
    lib.b200q_strerror.restype = ctypes.c_char_p
    lib.b200q_last_cuda_error.restype = ctypes.c_int
    lib.b200q_version.restype = ctypes.c_int
    lib.b200q_debug_decode_plan.argtypes = [ctypes.POINTER(Layer), ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)]
    lib.b200q_debug_decode_plan.restype = ctypes.c_int
    lib.b200q_debug_set_option.argtypes = [ctypes.c_char_p, ctypes.c_double]
    lib.b200q_debug_set_option.restype = ctypes.c_int
    lib.b200q_debug_set_chain_timeline.argtypes = [P]
    lib.b200q_debug_set_chain_timeline.restype = None
    lib.b200q_debug_set_timeline.argtypes = [P, SZ]
    lib.b200q_debug_set_timeline.restype = None
    return lib


lib = _load()


def check(status: int, what: str = "b200q"):
    """Map status codes to the exception types the reference raises (SURVEY §8b errors row)."""
    if status == B200Q_OK:
        return
    msg = f"{what}: {lib.b200q_strerror(status).decode()} (status {status})"
    if status == -6:
        msg += f", cudaError={lib.b200q_last_cuda_error()}"
        raise RuntimeError(msg)
    if status in (-2, -3, -4):
        raise ValueError(msg)
    raise RuntimeError(msg)
