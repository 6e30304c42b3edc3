"""Tensor-level wrappers over the C-ABI (``include/pnp_b200.h``).

Every function takes CUDA tensors, launches on ``torch.cuda.current_stream()`` and returns CUDA tensors;
nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.PnpError(f"{name}: expected a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    _lib.check_device(t.device)
    return t.contiguous()


def aligned_empty(nbytes: int, device, align: int = 1024) -> torch.Tensor:
    """uint8 CUDA buffer whose data_ptr is ``align``-byte aligned."""
    raw = torch.empty(nbytes + align, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % align
    return raw[off:off + nbytes]


# ------------------------------------------------------------------------------------------------
def psnr(x: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """``[N,...]`` vs ``[N,...]`` (or one shared gt) -> fp32 ``[N]`` on the device (env.py:120-125)."""
    if x.is_complex():
        x = x.real
    N = x.shape[0]
    x = _req(x.reshape(N, -1).float(), torch.float32, "x")
    gt = _req(gt.float(), torch.float32, "gt")
    HW = x.shape[1]
    if gt.numel() == N * HW:
        stride = HW
    elif gt.numel() == HW:
        stride = 0
    else:
        raise RuntimeError(f"psnr: gt has {gt.numel()} elements, expected {N * HW} or {HW}")
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    check(_lib.lib().pnp_psnr(x.data_ptr(), gt.data_ptr(), stride, out.data_ptr(), N, HW, _lib.stream_ptr()), "pnp_psnr")
    return out


def fft2c(x: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """Centred orthonormal 2-D (i)FFT over the last two dims (transformations.py:6-19)."""
    if not x.is_complex():
        x = torch.complex(x.float(), torch.zeros_like(x, dtype=torch.float32))
    x = _req(x, torch.complex64, "x")
    H, W = x.shape[-2:]
    B = x.numel() // (H * W)
    out = torch.empty_like(x)
    check(_lib.lib().pnp_fft2c(x.data_ptr(), out.data_ptr(), B, H, W, int(inverse), _lib.stream_ptr()), "pnp_fft2c")
    return out


def residual_real(z: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """``Re(z - u)`` as fp32 (env.py:85-86)."""
    z = _req(z, torch.complex64, "z")
    u = _req(u, torch.complex64, "u")
    v = torch.empty(z.shape, dtype=torch.float32, device=z.device)
    check(_lib.lib().pnp_residual_real(z.data_ptr(), u.data_ptr(), v.data_ptr(), z.numel(), _lib.stream_ptr()),
          "pnp_residual_real")
    return v


class MaskKindProbe:
    """Host-side view of the mask-structure flag that ``pnp_prox_prepare`` leaves on the device, WITHOUT a synchronisation:
    an asynchronous 4-byte copy into pinned memory plus an event.  ``get()`` returns -1 until the copy has landed, then
    1 (every mask depends on the column index only: row kernel) or 0 (general masks: cluster kernel) - the ``kind``
    argument of ``pnp_prox_dual_prepared_kind`` / ``pnp_step_prepared_kind`` (-1 launches both kernels)."""

    def __init__(self, maskp: torch.Tensor, mstride: int, B: int, H: int, W: int):
        self.host = torch.full((1,), -1, dtype=torch.int32).pin_memory()
        check(_lib.lib().pnp_prox_prepared_kind_async(maskp.data_ptr(), mstride, B, H, W, self.host.data_ptr(),
                                                      _lib.stream_ptr()), "pnp_prox_prepared_kind_async")
        self.event = torch.cuda.Event()
        self.event.record()
        self.kind = -1

    def get(self) -> int:
        if self.kind < 0 and not torch.cuda.is_current_stream_capturing() and self.event.query():
            self.kind = 1 if int(self.host[0]) != 0 else 0
        return self.kind


class ProxPrepared:
    """Trajectory constants of the prox step, prepared once (``pnp_prox_prepare``): see include/pnp_b200.h.

    ``ProxPrepared(y0, mask)`` then ``prep.prox_dual(x, u, mu)`` per iteration; same results as ``prox_dual``.
    Only for shapes with a prepared path (``supported(H, W)``: 256x256)."""

    @staticmethod
    def supported(H: int, W: int) -> bool:
        return bool(_lib.lib().pnp_prox_prepared_supported(H, W))

    def __init__(self, y0, mask):
        import ctypes as C
        y0 = _req(y0, torch.complex64, "y0")
        H, W = y0.shape[-2:]
        B = y0.numel() // (H * W)
        if mask.dtype == torch.bool:
            mask = mask.contiguous().view(torch.uint8)
        mask = _req(mask, torch.uint8, "mask")
        if mask.numel() == B * H * W:
            self.mstride = H * W
        elif mask.numel() == H * W:
            self.mstride = 0
        else:
            raise IndexError(f"mask with {mask.numel()} elements does not match y0 {tuple(y0.shape)}")
        self.B, self.H, self.W = B, H, W
        l = _lib.lib()
        n_y, n_m = C.c_size_t(0), C.c_size_t(0)
        check(l.pnp_prox_prepared_bytes(B, H, W, C.byref(n_y), C.byref(n_m)), "pnp_prox_prepared_bytes")
        self.y0p = torch.empty(n_y.value // 8, dtype=torch.complex64, device=y0.device)
        self.maskp = torch.empty(n_m.value, dtype=torch.uint8, device=y0.device)
        check(l.pnp_prox_prepare(y0.data_ptr(), mask.data_ptr(), self.mstride, self.y0p.data_ptr(), self.maskp.data_ptr(),
                                 B, H, W, _lib.stream_ptr()), "pnp_prox_prepare")
        self.probe = MaskKindProbe(self.maskp, self.mstride, B, H, W)

    @property
    def column_only(self) -> bool:
        """Did the preparation find every mask of the batch to depend on the column index only? (synchronises)"""
        nb = self.B if self.mstride else 1                   # layout documented in csrc/fftprox.cu (prox_prepared_bytes)
        stride = 32 if (self.H == 256 and self.W == 256) else self.W
        off = ((nb * self.H * self.W + 15) // 16 * 16 + nb * stride + 15) // 16 * 16
        return bool(self.maskp[off:off + 4].view(torch.int32).item() != 0)

    def prox_dual(self, x, u, mu, want_v: bool = True, out=None, kind: int | None = None):
        """``kind``: None = use the host-side hint once it has landed; -1 forces the both-kernels launch."""
        x = _req(x, torch.float32, "x")
        u = _req(u, torch.complex64, "u")
        B = self.B
        mu = _req(mu.reshape(-1).float(), torch.float32, "mu")
        if mu.numel() not in (1, B):
            raise RuntimeError(f"mu must have 1 or {B} elements, got {mu.numel()}")
        if out is None:
            z, un = torch.empty_like(u), torch.empty_like(u)
            v = torch.empty_like(x) if want_v else None
        else:
            z, un, v = out
        check(_lib.lib().pnp_prox_dual_prepared_kind(x.data_ptr(), u.data_ptr(), self.y0p.data_ptr(), self.maskp.data_ptr(),
                                                     self.mstride, mu.data_ptr(), 0 if mu.numel() == 1 else 1, z.data_ptr(),
                                                     un.data_ptr(), v.data_ptr() if v is not None else None, B, self.H,
                                                     self.W, self.probe.get() if kind is None else kind, _lib.stream_ptr()),
              "pnp_prox_dual_prepared_kind")
        return z, un, v


def prox_dual(x, u, y0, mask, mu, want_v: bool = True, out=None, workspace=None):
    """env.py:87-93.  x fp32 ``[B,1,H,W]``; u, y0 c64; mask bool/uint8 ``[B or 1,1,H,W]``; mu fp32 ``[1]`` or ``[B]``.

    Returns ``(z, u_new, v_next)`` as fresh tensors unless ``out=(z, u_new, v)`` is given.
    """
    x = _req(x, torch.float32, "x")
    u = _req(u, torch.complex64, "u")
    y0 = _req(y0, torch.complex64, "y0")
    H, W = x.shape[-2:]
    B = x.numel() // (H * W)
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    mask = _req(mask, torch.uint8, "mask")
    if mask.numel() == B * H * W:
        mstride = H * W
    elif mask.numel() == H * W:
        mstride = 0
    else:
        raise IndexError(f"mask with {mask.numel()} elements does not match x {tuple(x.shape)}")
    mu = _req(mu.reshape(-1).float(), torch.float32, "mu")
    if mu.numel() == 1:
        mu_stride = 0
    elif mu.numel() == B:
        mu_stride = 1
    else:
        raise RuntimeError(f"mu must have 1 or {B} elements, got {mu.numel()}")
    if out is None:
        z = torch.empty_like(u)
        un = torch.empty_like(u)
        v = torch.empty_like(x) if want_v else None
    else:
        z, un, v = out
    l = _lib.lib()
    if workspace is None:
        workspace = torch.empty(l.pnp_prox_workspace_bytes(B, H, W), dtype=torch.uint8, device=x.device)
    check(l.pnp_prox_dual(x.data_ptr(), u.data_ptr(), y0.data_ptr(), mask.data_ptr(), mstride, mu.data_ptr(), mu_stride,
                          z.data_ptr(), un.data_ptr(), v.data_ptr() if v is not None else None, workspace.data_ptr(),
                          B, H, W, _lib.stream_ptr()), "pnp_prox_dual")
    return z, un, v


def conv3x3_bf16(in0: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, in1: torch.Tensor | None = None,
                 in1_half_res: bool = False):
    """NHWC bf16 3x3 conv + bias + LeakyReLU(0.2) on the tensor cores; ``in1`` = second concat segment.
    ``in1_half_res``: ``in1`` is ``[B,H/2,W/2,C1]`` and its x2 bilinear upsample (align_corners=True) is the segment
    (the fused first conv of an ``up`` block, reference noise.py:39,59)."""
    in0 = _req(in0, torch.bfloat16, "in0")
    B, H, W, C0 = in0.shape
    C1 = 0
    if in1 is not None:
        in1 = _req(in1, torch.bfloat16, "in1")
        C1 = in1.shape[-1]
        if in1_half_res and tuple(in1.shape[1:3]) != (H // 2, W // 2):
            raise ValueError(f"half-resolution segment must be {(H // 2, W // 2)}, got {tuple(in1.shape[1:3])}")
    weight = _req(weight, torch.float32, "weight")
    bias = _req(bias, torch.float32, "bias")
    Cout = weight.shape[0]
    assert weight.shape == (Cout, C0 + C1, 3, 3)
    l = _lib.lib()
    out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=in0.device)
    scratch = aligned_empty(l.pnp_conv3x3_packed_bytes(C0 + C1, Cout), in0.device)
    fn = l.pnp_conv3x3_ups_bf16 if in1_half_res else l.pnp_conv3x3_bf16
    check(fn(in0.data_ptr(), C0, in1.data_ptr() if in1 is not None else None, C1, weight.data_ptr(),
             bias.data_ptr(), out.data_ptr(), scratch.data_ptr(), B, H, W, Cout, _lib.stream_ptr()),
          "pnp_conv3x3_ups_bf16" if in1_half_res else "pnp_conv3x3_bf16")
    return out


# ------------------------------------------------------------------------------------------------
_UNET_BLOCKS = (("inc.conv", 2, 32), ("down1.mpconv.1", 32, 64), ("down2.mpconv.1", 64, 128), ("down3.mpconv.1", 128, 256),
                ("down4.mpconv.1", 256, 512), ("up1.conv", 768, 256), ("up2.conv", 384, 128), ("up3.conv", 192, 64),
                ("up4.conv", 96, 32))


def unet_state_dict_shapes():
    """(key, shape) of the 56 tensors of the reference ``UNet(2, 1)`` state_dict in registration order."""
    out = []
    for blk, cin, cout in _UNET_BLOCKS:
        for i in range(3):
            out.append((f"{blk}.conv-{i}.conv2d.weight", (cout, cin if i == 0 else cout, 3, 3)))
            out.append((f"{blk}.conv-{i}.conv2d.bias", (cout,)))
    out += [("outc.conv.weight", (1, 32, 1, 1)), ("outc.conv.bias", (1,))]
    return out


def unflatten_state_dict(flat: torch.Tensor):
    """Inverse of ``flatten_state_dict``: flat fp32 vector -> reference-format state_dict."""
    from collections import OrderedDict
    flat = flat.detach().reshape(-1).to(torch.float32).cpu()
    sd, off = OrderedDict(), 0
    for k, shp in unet_state_dict_shapes():
        n = 1
        for d in shp:
            n *= d
        sd[k] = flat[off:off + n].reshape(shp).clone()
        off += n
    if off != flat.numel():
        raise ValueError(f"flat parameter vector has {flat.numel()} entries, the U-Net has {off}")
    return sd


def flatten_state_dict(sd) -> torch.Tensor:
    """Reference ``UNet(2,1)`` state_dict (56 tensors, noise.py:101-113) -> flat fp32 vector in registration order."""
    keys = [k for k, _ in unet_state_dict_shapes()]
    missing = [k for k in keys if k not in sd]
    if missing:
        raise KeyError(f"state_dict is missing U-Net tensors: {missing[:4]}{'...' if len(missing) > 4 else ''}")
    return torch.cat([sd[k].detach().reshape(-1).to(torch.float32).cpu() for k in keys])


class UNetPlan:
    """Launch plan of the denoiser for one ``(B, H, W)`` (tensor maps, activation workspace)."""

    def __init__(self, packed: torch.Tensor, B: int, H: int, W: int):
        l = _lib.lib()
        self.B, self.H, self.W = B, H, W
        self.packed = packed
        nbytes = l.pnp_unet_workspace_bytes(B, H, W)
        self.workspace = aligned_empty(nbytes, packed.device)
        h = C.c_void_p()
        check(l.pnp_unet_plan_create(C.byref(h), packed.data_ptr(), self.workspace.data_ptr(), nbytes, B, H, W),
              "pnp_unet_plan_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().pnp_unet_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def forward(self, v: torch.Tensor, sigma: torch.Tensor, out: torch.Tensor | None = None, preclamp: bool = False):
        v = _req(v, torch.float32, "v")
        sigma = _req(sigma.reshape(-1).float(), torch.float32, "sigma")
        assert v.numel() == self.B * self.H * self.W and sigma.numel() == self.B
        if out is None:
            out = torch.empty_like(v)
        pre = torch.empty_like(v) if preclamp else None
        check(_lib.lib().pnp_unet_forward(self.handle, v.data_ptr(), sigma.data_ptr(), out.data_ptr(),
                                          pre.data_ptr() if pre is not None else None, _lib.stream_ptr()),
              "pnp_unet_forward")
        return (out, pre) if preclamp else out

    def activation(self, name: str) -> torch.Tensor:
        """NHWC bf16 view of a named intermediate (valid after ``forward``; later layers may reuse buffers)."""
        off, c, h, w = C.c_size_t(), C.c_int(), C.c_int(), C.c_int()
        rc = _lib.lib().pnp_unet_plan_tensor(self.handle, name.encode(), C.byref(off), C.byref(c), C.byref(h), C.byref(w))
        if rc != 0:
            raise KeyError(name)
        n = self.B * h.value * w.value * c.value
        return self.workspace[off.value:off.value + 2 * n].view(torch.bfloat16).view(self.B, h.value, w.value, c.value)


def pack_unet_weights(flat: torch.Tensor) -> torch.Tensor:
    """fp32 flat parameter vector (CUDA) -> packed bf16 tensor-core blobs (+ fp32 copy)."""
    l = _lib.lib()
    flat = _req(flat, torch.float32, "flat")
    if flat.numel() != l.pnp_unet_num_params():
        raise RuntimeError(f"expected {l.pnp_unet_num_params()} U-Net parameters, got {flat.numel()}")
    packed = aligned_empty(l.pnp_unet_packed_bytes(), flat.device)
    packed.zero_()
    check(l.pnp_unet_pack_weights(flat.data_ptr(), packed.data_ptr(), _lib.stream_ptr()), "pnp_unet_pack_weights")
    return packed
