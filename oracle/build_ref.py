#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/ from the reference's own sources where they lie under
/root/reference (nothing is copied into the repository; oracle/_ref/ is git-ignored but travels to the GPU box):

  libmultibox_ref.so   operator/multibox_{prior,target,detection}.cc compiled in place behind oracle/shim/mxnet_shim.h
                       (the function templates MultiBox*Forward<cpu> are the reference's code, verbatim; the
                       Forward() glue from the -inl.h headers is restated in oracle/shim/ref_*.cc);
  cpu_nms_ref*.so      cython/cpu_nms.pyx with the 3-token numpy-2 / Cython-3 patch of SURVEY.md section 8c
                       (np.int_t -> np.intp_t, dtype=np.int -> np.intp, `np.float thresh` -> `double thresh`),
                       applied on the fly to a scratch copy under oracle/_ref/;
  bbox_ref*.so         cython/bbox.pyx (bbox_overlaps_cython) with `np.float` -> `np.float64`;
  libgpu_nms_ref.so    cython/nms_kernel.cu + gpu_nms.hpp, UNMODIFIED, compiled with nvcc for sm_100a: the reference's own
                       GPU NMS (`_nms`: cudaMalloc, H2D, full N x N/64 mask, D2H, host sweep), used only as the same-box
                       speed comparator of bench.py --workload nms.

The reference's own build (MXNet's make/cmake with the operators dropped into src/operator/contrib) cannot be run:
MXNet is neither vendored nor installable here.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DSPNET_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w"]


def build_operators():
    objs = []
    for op in ("prior", "target", "detection"):
        obj = os.path.join(OUT, "ref_%s.o" % op)
        subprocess.check_call(["g++"] + CXXFLAGS + ["-I", os.path.join(HERE, "shim"), "-I", os.path.join(HERE, "shim", "operator"),
                                                  "-DREF_SOURCE(f)=<%s/operator/f>" % REF, "-c",
                                                  os.path.join(HERE, "shim", "ref_%s.cc" % op), "-o", obj])
        objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-o", os.path.join(OUT, "libmultibox_ref.so")] + objs)
    for o in objs:
        os.unlink(o)


def build_cpu_nms():
    import numpy
    src = open(os.path.join(REF, "cython", "cpu_nms.pyx")).read()
    patched = src.replace("np.int_t", "np.intp_t").replace("dtype=np.int)", "dtype=np.intp)").replace(
        "np.float thresh", "double thresh")
    assert patched != src
    pyx = os.path.join(OUT, "cpu_nms_ref.pyx")
    with open(pyx, "w") as f:
        f.write(patched)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--fast-fail", pyx, "-o", os.path.join(OUT, "cpu_nms_ref.c")])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
                           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", "-I", sysconfig.get_paths()["include"],
                           "-I", numpy.get_include(), os.path.join(OUT, "cpu_nms_ref.c"), "-o",
                           os.path.join(OUT, "cpu_nms_ref" + ext)])
    os.unlink(os.path.join(OUT, "cpu_nms_ref.c"))
    os.unlink(pyx)


def build_bbox():
    """cython/bbox.pyx with the numpy-2 patch `np.float` -> `np.float64` (the alias was removed from numpy)."""
    import numpy
    src = open(os.path.join(REF, "cython", "bbox.pyx")).read()
    patched = src.replace("DTYPE = np.float\n", "DTYPE = np.float64\n").replace("ctypedef np.float_t DTYPE_t", "ctypedef np.float64_t DTYPE_t")
    assert patched != src
    pyx = os.path.join(OUT, "bbox_ref.pyx")
    with open(pyx, "w") as f:
        f.write(patched)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--fast-fail", pyx, "-o", os.path.join(OUT, "bbox_ref.c")])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-w",
                           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION", "-I", sysconfig.get_paths()["include"],
                           "-I", numpy.get_include(), os.path.join(OUT, "bbox_ref.c"), "-o",
                           os.path.join(OUT, "bbox_ref" + ext)])
    os.unlink(os.path.join(OUT, "bbox_ref.c"))
    os.unlink(pyx)


def build_gpu_nms():
    """The reference's nms_kernel.cu as it is (the file includes "gpu_nms.hpp" from its own directory)."""
    subprocess.check_call(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                           "-cudart", "static", "-w", os.path.join(REF, "cython", "nms_kernel.cu"), "-o",
                           os.path.join(OUT, "libgpu_nms_ref.so")])


def main():
    if not os.path.isdir(os.path.join(REF, "operator")):
        print("build_ref: %s not present, keeping the prebuilt oracle/_ref/ (if any)" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    build_operators()
    try:
        build_cpu_nms()
    except Exception as e:  # Cython missing etc.: the operators are the important part
        print("build_ref: cpu_nms not built (%r)" % (e,))
    try:
        build_bbox()
    except Exception as e:
        print("build_ref: bbox not built (%r)" % (e,))
    try:
        build_gpu_nms()
    except Exception as e:
        print("build_ref: gpu_nms comparator not built (%r)" % (e,))
    print("build_ref: ok ->", sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
