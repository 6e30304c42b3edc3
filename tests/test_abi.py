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


def _oracle_imports(path):
    """(function name or None, line) of every statement in ``path`` that imports something from ``oracle``."""
    import ast
    tree = ast.parse(open(path).read())
    found = []

    def visit(node, fn):
        for child in ast.iter_child_nodes(node):
            name = child.name if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef)) else fn
            if isinstance(child, ast.ImportFrom) and (child.module or "").split(".")[0] == "oracle":
                found.append((fn, child.lineno))
            elif isinstance(child, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in child.names):
                found.append((fn, child.lineno))
            visit(child, name)

    visit(tree, None)
    return found


def test_product_path_never_imports_the_oracle():
    """The oracle is the checker: the package must not import it anywhere, and bench.py only inside its CPU leg."""
    pkg = os.path.join(ROOT, "dt4image_restoration_b200")
    for f in sorted(os.listdir(pkg)):
        if f.endswith(".py"):
            assert _oracle_imports(os.path.join(pkg, f)) == [], f
    fns = {fn for fn, _ in _oracle_imports(os.path.join(ROOT, "bench.py"))}
    assert fns == {"cpu_leg"}, fns


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/pnp_b200.h compiles as C99 (-pedantic) and a C host links the library directly (examples/host_query.c)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("gcc not available")
    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "host_query")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "host_query.c"), "-o", exe, "-L", libdir, "-lpnp_b200",
                           f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([exe], text=True)
    assert "abi 1" in out and "unet params 11773857" in out and "prepared_supported 1" in out


def test_device_guard_message_is_explicit():
    """ADVICE (round 1): nothing in the package may run on a device other than the one pnp_init prepared.  The guard lives
    in _lib.check_device; here only its wiring is checked (no GPU): ops._req and PnPEngine call it."""
    import inspect
    from dt4image_restoration_b200 import _lib, engine, ops
    assert "check_device" in inspect.getsource(ops._req)
    assert "check_device" in inspect.getsource(engine.PnPEngine.__init__)
    assert "one process per GPU" in inspect.getsource(_lib.check_device).lower() or "one process" in _lib.check_device.__doc__.lower()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_c_host_launches_kernels_and_matches_plain_c_arithmetic(tmp_path):
    """examples/host_prox.c: a C99 host (CUDA runtime + the C-ABI, no Python, no torch) runs pnp_prox_dual, pnp_fft2c and
    pnp_psnr at 32x32 (radix kernels) and 18x24 (dense-DFT path) and checks them against its own O(N^4) double-precision
    restatement of the reference's centred transforms / masked solve / dual update / PSNR."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    cuda = next((d for d in ("/usr/local/cuda", "/usr/local/cuda-12.9") if os.path.exists(os.path.join(d, "include", "cuda_runtime_api.h"))), None)
    if gcc is None or cuda is None:
        pytest.skip("gcc or the CUDA runtime headers are not available")
    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "host_prox")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(cuda, "include"), os.path.join(ROOT, "examples", "host_prox.c"), "-o", exe,
                           "-L", libdir, "-lpnp_b200", f"-Wl,-rpath,{libdir}", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm"])
    out = subprocess.run([exe], text=True, capture_output=True)
    assert out.returncode == 0 and "host_prox: ok" in out.stdout, out.stdout + out.stderr
