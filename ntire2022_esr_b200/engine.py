"""Python mirror of the C ABI: one `Engine` = one esr_handle bound to one CUDA device.

`Engine` replaces the reference's module construction + `load_state_dict(strict=True)` + `.to(device)`
(test_demo.py:17-30,52-58,150-157,336-340) and `model(img_lq)` (test_demo.py:367).  PyTorch is only
used for device memory and the current stream; all compute happens in libesr_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Mapping, Optional

import numpy as np

from . import _cabi
from ._cabi import EsrError, lib

ARCHS = {"imdn": _cabi.ARCH_IMDN, "rfdn": _cabi.ARCH_RFDN, "rlfn": _cabi.ARCH_RLFN, "bsrn": _cabi.ARCH_BSRN,
         "rfdn_pruned": _cabi.ARCH_RFDN_PRUNED, "fmen": _cabi.ARCH_FMEN}


def _as_f32_numpy(v) -> np.ndarray:
    if hasattr(v, "detach"):  # torch tensor
        v = v.detach().to("cpu").float().numpy()
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32))


class Engine:
    """esr_handle wrapper.  device=-1 gives a host-only handle (weights can be loaded and checked,
    every compute call raises: the engine has no CPU path)."""

    def __init__(self, arch: str, device: int = 0, nf: int = 0, nblocks: int = 0):
        if arch not in ARCHS:
            raise NotImplementedError(f"architecture {arch!r} is not on the accelerated path")
        self.arch = arch
        self.device = device
        self._h = ctypes.c_void_p()
        rc = lib.esr_create(ctypes.byref(self._h), ARCHS[arch], nf, nblocks, device)
        if rc != _cabi.OK:
            self._h = ctypes.c_void_p()
            msg = {_cabi.E_NOGPU: f"CUDA device {device} is not a usable sm_100 GPU (no CPU fallback exists)",
                   _cabi.E_INVALID: "invalid architecture / width / depth"}.get(rc, "esr_create failed")
            raise EsrError(rc, msg)
        self._finalized = False
        self._ws = None

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.esr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != _cabi.OK:
            raise EsrError(rc, (lib.esr_last_error(self._h) or b"").decode())

    # -- weights ----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Mapping[str, object]):
        """strict=True semantics: missing / unexpected / mis-shaped entries raise at finalize."""
        for name, v in state_dict.items():
            a = _as_f32_numpy(v)
            shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
            self._check(lib.esr_load_weights(self._h, name.encode(), a.ctypes.data_as(ctypes.c_void_p), shape, a.ndim))
        self._check(lib.esr_finalize(self._h))
        self._finalized = True
        return self

    def set_option(self, key: str, value: int):
        self._check(lib.esr_set_option(self._h, key.encode(), int(value)))

    # -- queries ----------------------------------------------------------------------------------
    def workspace_bytes(self, B: int, H: int, W: int, dtype: int) -> int:
        n = lib.esr_workspace_bytes(self._h, B, H, W, dtype)
        if n == 0:
            raise EsrError(_cabi.E_INVALID, (lib.esr_last_error(self._h) or b"").decode())
        return n

    def launch_names(self, B: int, H: int, W: int, dtype: int):
        n = lib.esr_launch_count(self._h, B, H, W, dtype)
        return [lib.esr_launch_name(self._h, B, H, W, dtype, i).decode() for i in range(n)]

    def launch_flops(self, B: int, H: int, W: int, dtype: int):
        n = lib.esr_launch_count(self._h, B, H, W, dtype)
        return [lib.esr_launch_flops(self._h, B, H, W, dtype, i) for i in range(n)]

    def profile_launches(self, x, out, reps: int = 20):
        """[(launch name, algorithmic FLOPs, mean ms)] for one forward of the CUDA tensor x."""
        import torch

        B, _, H, W = x.shape
        dt = _cabi.DTYPE_F16 if x.dtype == torch.float16 else _cabi.DTYPE_F32
        self.forward(x, out=out)  # sizes the workspace
        names = self.launch_names(B, H, W, dt)
        flops = self.launch_flops(B, H, W, dt)
        ms = (ctypes.c_float * len(names))()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        n = lib.esr_profile_launches(self._h, x.data_ptr(), out.data_ptr(), B, H, W, dt, self._ws.data_ptr(),
                                     self._ws.numel(), reps, ms, len(names), stream)
        if n < 0:
            self._check(n)
        return [(names[i], flops[i], float(ms[i])) for i in range(n)]

    def debug_timeline(self, n_launches: int):
        buf = (ctypes.c_longlong * (128 * n_launches))()
        self._check(lib.esr_debug_timeline(self._h, buf, n_launches))
        return np.frombuffer(buf, dtype=np.int64).reshape(n_launches, 4, 32).copy()

    @staticmethod
    def _check_out(out, shape, dtype, device):
        """A caller-supplied output goes to the engine as a raw pointer: refuse anything the kernels would overrun."""
        if (not hasattr(out, "data_ptr") or tuple(out.shape) != tuple(shape) or out.dtype != dtype or out.device != device
                or not out.is_contiguous() or out.data_ptr() % 16):
            raise EsrError(_cabi.E_INVALID, f"`out` must be a contiguous, 16-byte aligned {dtype} tensor of shape {tuple(shape)} on {device}")

    # -- compute ----------------------------------------------------------------------------------
    def forward(self, x, out=None):
        """x: CUDA torch tensor (B,3,H,W) fp32 or fp16 on this engine's device -> (B,3,4H,4W)."""
        import torch

        if not x.is_cuda:
            raise EsrError(_cabi.E_NOGPU, "input is not a CUDA tensor; the engine has no CPU fallback")
        if x.device.index != self.device:
            raise EsrError(_cabi.E_INVALID, f"input on {x.device}, engine bound to cuda:{self.device}")
        if x.dim() != 4 or x.shape[1] != 3:
            raise EsrError(_cabi.E_INVALID, f"expected (B,3,H,W), got {tuple(x.shape)}")
        if x.dtype == torch.float32:
            dt = _cabi.DTYPE_F32
        elif x.dtype == torch.float16:
            dt = _cabi.DTYPE_F16
        else:
            raise EsrError(_cabi.E_INVALID, f"unsupported dtype {x.dtype}")
        x = x.contiguous()
        if x.data_ptr() % 16:
            x = x.clone()
        B, _, H, W = x.shape
        need = self.workspace_bytes(B, H, W, dt)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        if out is None:
            out = torch.empty((B, 3, 4 * H, 4 * W), dtype=x.dtype, device=x.device)
        else:
            self._check_out(out, (B, 3, 4 * H, 4 * W), x.dtype, x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        self._check(lib.esr_forward(self._h, x.data_ptr(), out.data_ptr(), B, H, W, dt, self._ws.data_ptr(),
                                    self._ws.numel(), stream))
        return out

    def forward_uint8(self, img, data_range: float, half: bool = True, out=None):
        """img: CUDA uint8 tensor (B,H,W,3) or (H,W,3) -> uint8 (B,4H,4W,3) / (4H,4W,3).

        One call for the reference's uint2tensor4 -> forward -> tensor2uint (test_demo.py:423-434)."""
        import torch

        if not img.is_cuda or img.dtype != torch.uint8:
            raise EsrError(_cabi.E_INVALID, "expected a CUDA uint8 tensor (the engine has no CPU fallback)")
        squeeze = img.dim() == 3
        if squeeze:
            img = img.unsqueeze(0)
        if img.dim() != 4 or img.shape[3] != 3:
            raise EsrError(_cabi.E_INVALID, f"expected (B,H,W,3), got {tuple(img.shape)}")
        img = img.contiguous()
        B, H, W, _ = img.shape
        dt = _cabi.DTYPE_F16 if half else _cabi.DTYPE_F32
        need = lib.esr_workspace_bytes_u8(self._h, B, H, W, dt)
        if need == 0:
            raise EsrError(_cabi.E_INVALID, (lib.esr_last_error(self._h) or b"").decode())
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=img.device)
        if out is None:
            out = torch.empty((B, 4 * H, 4 * W, 3), dtype=torch.uint8, device=img.device)
        else:
            self._check_out(out, (B, 4 * H, 4 * W, 3), torch.uint8, img.device)
        stream = torch.cuda.current_stream(img.device).cuda_stream
        self._check(lib.esr_forward_u8(self._h, img.data_ptr(), out.data_ptr(), B, H, W, float(data_range), dt,
                                       self._ws.data_ptr(), self._ws.numel(), stream))
        return out[0] if squeeze else out

    def forward_host_uint8(self, img: np.ndarray, data_range: float, half: bool = True) -> np.ndarray:
        """Host uint8 (B,H,W,3) / (H,W,3) in, host uint8 out (esr_forward_host_u8)."""
        img = np.ascontiguousarray(img)
        if img.dtype != np.uint8:
            raise EsrError(_cabi.E_INVALID, f"unsupported dtype {img.dtype}")
        squeeze = img.ndim == 3
        if squeeze:
            img = img[None]
        if img.ndim != 4 or img.shape[3] != 3:
            raise EsrError(_cabi.E_INVALID, f"expected (B,H,W,3), got {img.shape}")
        B, H, W, _ = img.shape
        out = np.empty((B, 4 * H, 4 * W, 3), dtype=np.uint8)
        self._check(lib.esr_forward_host_u8(self._h, img.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                                            B, H, W, float(data_range), _cabi.DTYPE_F16 if half else _cabi.DTYPE_F32))
        return out[0] if squeeze else out

    def forward_host_u8_async_ptr(self, in_ptr: int, out_ptr: int, B: int, H: int, W: int, data_range: float, dt: int) -> int:
        t = ctypes.c_longlong(-1)
        self._check(lib.esr_forward_host_u8_async(self._h, in_ptr, out_ptr, B, H, W, float(data_range), dt, ctypes.byref(t)))
        return t.value

    def forward_host(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """x: host array (B,3,H,W) float32 / float16; copies in, runs, copies out, synchronises."""
        x = np.ascontiguousarray(x)
        if x.dtype == np.float32:
            dt = _cabi.DTYPE_F32
        elif x.dtype == np.float16:
            dt = _cabi.DTYPE_F16
        else:
            raise EsrError(_cabi.E_INVALID, f"unsupported dtype {x.dtype}")
        if x.ndim != 4 or x.shape[1] != 3:
            raise EsrError(_cabi.E_INVALID, f"expected (B,3,H,W), got {x.shape}")
        B, _, H, W = x.shape
        if out is None:
            out = np.empty((B, 3, 4 * H, 4 * W), dtype=x.dtype)
        elif (not isinstance(out, np.ndarray) or out.shape != (B, 3, 4 * H, 4 * W) or out.dtype != x.dtype
              or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]):
            raise EsrError(_cabi.E_INVALID, f"`out` must be a writeable C-contiguous {x.dtype} array of shape {(B, 3, 4 * H, 4 * W)}")
        self._check(lib.esr_forward_host(self._h, x.ctypes.data_as(ctypes.c_void_p),
                                         out.ctypes.data_as(ctypes.c_void_p), B, H, W, dt))
        return out

    def forward_host_async_ptr(self, in_ptr: int, out_ptr: int, B: int, H: int, W: int, dt: int) -> int:
        """Queue one request on pinned host buffers; returns its ticket (see esr_forward_host_async)."""
        t = ctypes.c_longlong(-1)
        self._check(lib.esr_forward_host_async(self._h, in_ptr, out_ptr, B, H, W, dt, ctypes.byref(t)))
        return t.value

    def host_wait(self, ticket: int = -1):
        self._check(lib.esr_host_wait(self._h, ticket))

    def forward_host_ptr(self, in_ptr: int, out_ptr: int, B: int, H: int, W: int, dt: int):
        """Raw-pointer variant (pinned torch tensors): no numpy wrapping on the timed path."""
        self._check(lib.esr_forward_host(self._h, in_ptr, out_ptr, B, H, W, dt))
