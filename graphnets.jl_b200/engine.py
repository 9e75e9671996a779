"""Per-device engine: owns one gnb_ctx, binds it to torch's current CUDA stream.

PyTorch is plumbing here (device memory, streams, torch.distributed); every compute call goes
through the C ABI of libgnb200.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    def __init__(self, device):
        self.device = int(device)
        err = C.c_int(0)
        self.ctx = lib.gnb_ctx_create(self.device, C.byref(err))
        if not self.ctx:
            check(err.value if err.value else _lib.GNB_ERR_CUDA)
        self.torch_device = torch.device("cuda", self.device)

    def bind_stream(self):
        s = torch.cuda.current_stream(self.torch_device).cuda_stream
        check(lib.gnb_ctx_set_stream(self.ctx, C.c_void_p(s)))

    def sync(self):
        check(lib.gnb_sync(self.ctx))

    @property
    def launches(self):
        return int(lib.gnb_ctx_launch_count(self.ctx))

    def set_profiling(self, on):
        check(lib.gnb_ctx_set_profiling(self.ctx, 1 if on else 0))

    def read_profile(self):
        """{kernel kind: dict(launches, ms, alg_bytes, alg_flops)} since the last read."""
        arr = (_lib.ProfEntry * 64)()
        n = C.c_int(0)
        check(lib.gnb_ctx_profile_read(self.ctx, arr, 64, C.byref(n)))
        return {arr[i].name.decode(): dict(launches=arr[i].launches, ms=arr[i].ms, alg_bytes=arr[i].alg_bytes,
                                           alg_flops=arr[i].alg_flops) for i in range(min(n.value, 64))}

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.torch_device)

    def __del__(self):
        try:
            if self.ctx:
                lib.gnb_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass


_engines = {}


def get_engine(device=None):
    """Engine of a CUDA device (default: torch's current device).  Raises if no GPU is present -
    the product path never falls back to the CPU."""
    if device is None:
        if not torch.cuda.is_available():
            # let the library produce its own loud error
            device = 0
        else:
            device = torch.cuda.current_device()
    if isinstance(device, torch.device):
        device = device.index if device.index is not None else 0
    device = int(device)
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
