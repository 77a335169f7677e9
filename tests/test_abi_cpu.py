"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/climsim_b200.h
declares, reports errors instead of computing on a machine without a GPU, and the header and ctypes table agree."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "climsim_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from climsim_b200 import _lib, build
    build.build()
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table and header disagree"
    assert lib.csb_version() == 100
    assert lib.csb_strerror(-2) == b"no sm_100 CUDA device"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_create_fails_loudly():
    from climsim_b200 import MLPEngine, _lib
    with pytest.raises(_lib.CsbError) as e:
        MLPEngine.mlp_v1(dtype="fp32", max_batch=128)
    assert e.value.code == _lib.ENODEV


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "climsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_flat_blob_keras_conversion_roundtrip():
    import numpy as np
    from climsim_b200.engine import MLPEngine
    from oracle.models import MLPRef
    ref = MLPRef(units=(64, 32), seed=1)
    ws = [p.detach().numpy() for p in ref.params]
    flat = MLPEngine.keras_to_flat(ws)
    assert flat.size == ref.num_parameters()
    # the fused head is [W_lin | W_relu] column-wise
    w_head = flat[-(128 * 128 + 128):-128].reshape(128, 128)
    np.testing.assert_array_equal(w_head[:, :120], ws[-4])
    np.testing.assert_array_equal(w_head[:, 120:], ws[-2])


def test_header_is_plain_c_and_the_c_example_links(tmp_path):
    """The boundary is a C ABI: include/climsim_b200.h compiles as C99 with -Wall -Werror, and examples/ffi_demo.c (create, set
    parameters, forward, train step, optimizer -- from C, no Python, no torch) compiles and links against the shared library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    cuda_inc = "/usr/local/cuda/include"
    if gcc is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("needs gcc and the CUDA runtime headers")
    from climsim_b200 import build
    build.build()
    hdr_only = tmp_path / "hdr.c"
    hdr_only.write_text('#include "climsim_b200.h"\nint main(void) { return csb_version() == CSB_VERSION ? 0 : 1; }\n')
    inc = ["-I", os.path.join(ROOT, "include"), "-I", cuda_inc]
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", *inc, "-c", str(hdr_only), "-o", str(tmp_path / "hdr.o")], check=True)
    obj = tmp_path / "ffi_demo.o"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", *inc, "-c", os.path.join(ROOT, "examples", "ffi_demo.c"), "-o", str(obj)], check=True)
    lib_dir = os.path.join(ROOT, "climsim_b200")
    subprocess.run([gcc, str(obj), "-L", lib_dir, "-lclimsim_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-o", str(tmp_path / "ffi_demo")],
                   check=True)
