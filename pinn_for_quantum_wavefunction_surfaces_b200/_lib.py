"""ctypes binding of libpinn_b200.so (include/pinn_b200.h).  No fallback: a missing library or a
non-B200 device raises."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libpinn_b200.so"

EXPORTS = [
    "pinn_version", "pinn_theta_size", "pinn_theta_offsets", "pinn_create", "pinn_destroy", "pinn_last_error",
    "pinn_launch_count", "pinn_set_engine", "pinn_get_engine", "pinn_profile_begin", "pinn_profile_collect", "pinn_loss_fwd_bwd", "pinn_fields", "pinn_loss_fwd_bwd_host",
]


class PinnError(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, _LIB_NAME)


_lib = None
_lock = threading.Lock()


def lib():
    """Load the shared library (built by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise PinnError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % path)
        L = ctypes.CDLL(path)
        vp, i32, i64, u32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, ctypes.c_float
        L.pinn_version.restype = i32
        L.pinn_theta_size.restype = i32
        L.pinn_theta_offsets.argtypes = [ctypes.POINTER(i32)]
        L.pinn_theta_offsets.restype = None
        L.pinn_create.argtypes = [i32, ctypes.POINTER(vp)]
        L.pinn_create.restype = i32
        L.pinn_destroy.argtypes = [vp]
        L.pinn_destroy.restype = i32
        L.pinn_last_error.argtypes = [vp]
        L.pinn_last_error.restype = ctypes.c_char_p
        L.pinn_launch_count.argtypes = [vp]
        L.pinn_launch_count.restype = i64
        L.pinn_set_engine.argtypes = [vp, i32]
        L.pinn_set_engine.restype = i32
        L.pinn_get_engine.argtypes = [vp]
        L.pinn_get_engine.restype = i32
        L.pinn_profile_begin.argtypes = [vp]
        L.pinn_profile_begin.restype = i32
        L.pinn_profile_collect.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)]
        L.pinn_profile_collect.restype = i32
        L.pinn_loss_fwd_bwd.argtypes = [vp, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, u32, f32, vp, vp, vp, vp]
        L.pinn_loss_fwd_bwd.restype = i32
        L.pinn_fields.argtypes = [vp, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.pinn_fields.restype = i32
        L.pinn_loss_fwd_bwd_host.argtypes = [vp, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, u32, f32, vp, vp, vp]
        L.pinn_loss_fwd_bwd_host.restype = i32
        _lib = L
        return L


class Handle:
    """One per process/device (SURVEY 8b threading: calls are serialised by stream order)."""
    _cache = {}
    _cache_lock = threading.Lock()

    def __init__(self, device=0):
        self.L = lib()
        self.device = int(device)
        hp = ctypes.c_void_p()
        rc = self.L.pinn_create(self.device, ctypes.byref(hp))
        if rc != 0:
            raise PinnError("pinn_create(device=%d) failed (%d): %s"
                            % (device, rc, self.L.pinn_last_error(None).decode()))
        self.h = hp

    @classmethod
    def get(cls, device=0):
        device = int(device)
        with cls._cache_lock:
            if device not in cls._cache:
                cls._cache[device] = Handle(device)
            return cls._cache[device]

    def check(self, rc, what):
        if rc != 0:
            raise PinnError("%s failed (%d): %s" % (what, rc, self.L.pinn_last_error(self.h).decode()))

    def launch_count(self):
        return int(self.L.pinn_launch_count(self.h))

    ENGINES = {"ffma": 0, "tcgen05": 1}

    def set_engine(self, name):
        """'tcgen05' (default) or 'ffma': which implementation of the fused step kernel runs."""
        self.check(self.L.pinn_set_engine(self.h, self.ENGINES[name]), "pinn_set_engine")

    def get_engine(self):
        return {v: k for k, v in self.ENGINES.items()}[int(self.L.pinn_get_engine(self.h))]

    def profile_begin(self):
        self.check(self.L.pinn_profile_begin(self.h), "pinn_profile_begin")

    def profile_collect(self):
        """-> (summed step-kernel milliseconds, number of step-kernel launches) since profile_begin."""
        ms, k = ctypes.c_double(), ctypes.c_int()
        self.check(self.L.pinn_profile_collect(self.h, ctypes.byref(ms), ctypes.byref(k)), "pinn_profile_collect")
        return ms.value, k.value

    def close(self):
        if self.h:
            self.L.pinn_destroy(self.h)
            self.h = None
