"""The Cython side of the drop-in boundary, compiled for real: tests/binding/gpu_nms_dspmb.pyx is the reference's
cython/gpu_nms.pyx:16-31 with its extern block pointed at dspmb.h / libdspmb.so.  The CPU test
cythonizes, compiles and links it and imports the module (no compute without a GPU); the GPU test calls it like
detect/nms.py:18-21 does and compares with the reference's GPU rule on the oracle."""
import importlib.util
import os
import subprocess
import sys
import sysconfig

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def binding(tmp_path_factory):
    pytest.importorskip("Cython")
    lib = os.path.join(ROOT, "dspnet_b200", "libdspmb.so")
    if not os.path.exists(lib):
        pytest.skip("libdspmb.so not built")
    out = tmp_path_factory.mktemp("binding")
    pyx = os.path.join(ROOT, "tests", "binding", "gpu_nms_dspmb.pyx")
    c = str(out / "gpu_nms_dspmb.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--fast-fail", pyx, "-o", c])
    so = str(out / ("gpu_nms_dspmb" + sysconfig.get_config_var("EXT_SUFFIX")))
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                           "-I", sysconfig.get_paths()["include"], "-I", np.get_include(), "-I", os.path.join(ROOT, "include"),
                           c, "-o", so, lib, "-Wl,-rpath," + os.path.dirname(lib)])
    spec = importlib.util.spec_from_file_location("gpu_nms_dspmb", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_binding_compiles_links_and_imports(binding):
    assert callable(binding.gpu_nms)
    assert binding.gpu_nms(np.zeros((0, 5), np.float32), 0.5) == []  # no boxes: no device call


@pytest.mark.gpu
def test_binding_matches_the_gpu_nms_rule(binding, oracle, cuda):
    from dspnet_b200 import synth
    dets = synth.nms_boxes(77, 3000)
    got = [int(i) for i in binding.gpu_nms(dets, 0.45, 0)]
    assert got == oracle.cpu_nms(dets, 0.45, mode="gpu")
