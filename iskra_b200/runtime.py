"""One device context per grid: owns the iskb_ctx handle (particles, fields, solver in HBM)."""
import ctypes as C
import os

import numpy as np

from . import _lib as L


def default_device():
    return int(os.environ.get("LOCAL_RANK", "0"))


class Runtime:
    def __init__(self, device=None):
        self.lib = L.lib()
        h = L.vp()
        L.check(self.lib.iskb_create(default_device() if device is None else device, C.byref(h)))
        self.h = h
        self.grid = None
        self.n_ranks, self.rank = 1, 0

    def close(self):
        if self.h:
            self.lib.iskb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        L.check(self.lib.iskb_set_stream(self.h, L.vp(cuda_stream_ptr)))

    def use_torch_stream(self):
        """Run the library on torch's current stream so that torch.cuda.Event timings bracket its
        work.  torch's default stream has handle 0, which iskb_set_stream reads as "private
        stream": in that case a dedicated torch stream is created and made current first."""
        import torch
        if torch.cuda.current_stream().cuda_stream == 0:
            self._torch_stream = torch.cuda.Stream()
            torch.cuda.set_stream(self._torch_stream)
        self.set_stream(torch.cuda.current_stream().cuda_stream)

    def synchronize(self):
        L.check(self.lib.iskb_synchronize(self.h))

    def launch_count(self):
        v = L.i64()
        L.check(self.lib.iskb_launch_count(self.h, C.byref(v)))
        return v.value

    def profile(self, on):
        L.check(self.lib.iskb_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        ms, n = L.f64(), L.i64()
        L.check(self.lib.iskb_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def comm_init_torch(self):
        """Creates the NCCL communicator for rho; the unique id travels over torch.distributed."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            L.check(self.lib.iskb_comm_unique_id(L.ptr(idbuf)))
        t = torch.from_numpy(idbuf)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        idbuf = t.cpu().numpy().copy()
        L.check(self.lib.iskb_comm_init(self.h, world, rank, L.ptr(idbuf)))
        self.n_ranks, self.rank = world, rank

    # -- fields --------------------------------------------------------------------------------
    def fields(self, rho=True, phi=True, E=True):
        nx, ny = self.grid.n
        r = np.zeros((nx, ny), order="F") if rho else None
        p = np.zeros((nx, ny), order="F") if phi else None
        e = np.zeros((nx, ny, 3), order="F") if E else None
        L.check(self.lib.iskb_fields_download(self.h, L.ptr(r), L.ptr(p), L.ptr(e)))
        return r, p, e

    def set_fields(self, rho=None, phi=None, E=None):
        f = lambda a: None if a is None else np.asfortranarray(a, dtype=np.float64)
        rho, phi, E = f(rho), f(phi), f(E)
        L.check(self.lib.iskb_fields_upload(self.h, L.ptr(rho), L.ptr(phi), L.ptr(E)))

    def step(self, dt, n_steps=1):
        L.check(self.lib.iskb_step(self.h, float(dt), int(n_steps)))

    def set_after_push(self, mode_x, mode_y):
        L.check(self.lib.iskb_set_after_push(self.h, mode_x, mode_y))

    def set_sort_interval(self, k):
        L.check(self.lib.iskb_set_sort_interval(self.h, int(k)))

    def set_advance_path(self, path):
        """0: tile directory + incremental re-group (default); 1: per-warp windows re-grouped by radix sort."""
        L.check(self.lib.iskb_set_advance_path(self.h, int(path)))

    def set_lean(self, on):
        L.check(self.lib.iskb_set_lean(self.h, 1 if on else 0))

    def join(self):
        """Make the context stream wait for a field solve still in flight on the field stream."""
        L.check(self.lib.iskb_stream_join(self.h))

    def set_sort_policy(self, miss_threshold, max_interval, full_interval=0):
        L.check(self.lib.iskb_set_sort_policy(self.h, float(miss_threshold), int(max_interval)))
        L.check(self.lib.iskb_set_sort_full_interval(self.h, int(full_interval)))
