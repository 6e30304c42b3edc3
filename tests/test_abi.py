"""C-ABI surface: the library loads without a GPU and exports exactly what include/pnp_b200.h declares."""
import ctypes
import os
import re

from dt4image_restoration_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pnp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pnp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = header_symbols()
    assert len(names) >= 22
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pnp_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_lib.SIGNATURES) == names


def test_sizes_and_version_without_gpu():
    lib = _lib.load()
    assert lib.pnp_abi_version() == 1
    # 56 state_dict tensors of reference UNet(2,1): 11 773 857 parameters
    assert lib.pnp_unet_num_params() == 11773857
    assert lib.pnp_prox_workspace_bytes(3, 256, 256) == 3 * 256 * 256 * 9      # c64 scratch / y0T + maskT
    assert lib.pnp_unet_workspace_bytes(1, 256, 256) > 0
    assert lib.pnp_unet_packed_bytes() > 2 * 11773857
    assert lib.pnp_conv3x3_packed_bytes(64, 128) == 64 * 128 * 9 * 2


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "pnp_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S).lower()
    assert "at::" not in src and "c10::" not in src


def test_calls_fail_loudly_before_init():
    lib = _lib.load()
    if _lib._inited:
        return
    rc = lib.pnp_psnr(None, None, 0, None, 1, 16, None)
    assert rc != 0 and b"pnp_init" in lib.pnp_last_error()
