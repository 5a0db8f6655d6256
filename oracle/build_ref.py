"""oracle/build_ref.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

Compiles the reference's own batched-LU GPU path (MAGMA 2.10.0 magmablas kernels + cuBLAS) from
the sources where they lie under /root/reference into oracle/_ref/libmagma_ref.so, as the same-box
"kernel to beat" and as a second parity witness (pivots vs the reference's GPU results).

The reference ships z-masters; its d-sources are produced by its own precision generator
(tools/codegen.py, a standalone script -- this is not the reference's build system). Generated
files and objects live in a scratch directory under /tmp and are deleted; nothing from the
reference is copied into the repository. Recipe: SURVEY.md section 8c (47 sources, nvcc for sm_100).
"""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libmagma_ref.so")

SRC_D = {  # z-master -> generated with -p d
    "src": ["zgetrf_batched", "zgetrf_panel_batched", "zgetf2_batched", "zgetrs_batched", "zgesv_batched",
            "zgetrf_vbatched", "zgetrf_panel_vbatched", "zgetf2_vbatched"],
    "magmablas": ["zgetrf_batched_smallsq_noshfl.cu", "zgetf2_kernels.cu", "zgetf2_kernels_var.cu",
                  "zgesv_batched_small.cu", "zlaswp_batched.cu", "zlaswp_vbatched.cu", "ztrsm_batched_core.cpp",
                  "ztrsm_small_batched.cu", "ztrsv_batched.cu", "zgemv_batched.cpp", "zgemv_batched_core.cu",
                  "zgemv_batched_smallsq.cu", "zgemm_batched.cpp", "zgemm_batched_smallsq.cu",
                  "ztrsm_vbatched_core.cpp", "ztrsm_small_vbatched.cu", "zset_pointer.cu"],
    "interface_cuda": ["blas_z_v2.cpp"],
}
SRC_PLAIN = {
    "magmablas": ["getrf_setup_pivinfo.cu", "dgemm_batched_core.cu", "dgemm_vbatched_core.cu", "vbatched_aux.cu",
                  "vbatched_check.cu", "set_pointer.cu"],
    "control": ["get_batched_crossover.cpp", "get_ntcol.cpp", "get_batched_gemm_decision.cpp", "xerbla.cpp",
                "constants.cpp", "auxiliary.cpp"],
    "interface_cuda": ["interface.cpp", "alloc.cpp", "copy_v2.cpp", "error.cpp"],
}
# only needed to satisfy get_batched_crossover.cpp's geqrf tables (all four precisions)
SRC_ALLPREC = {"magmablas": ["zunm2r_batched_sm.cu"]}


def sh(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("failed: " + " ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return r.stdout


def main():
    if not os.path.isdir(REF):
        print("oracle/build_ref.py: /root/reference not present; keeping any prebuilt", LIB)
        return
    if os.path.exists(LIB) and "--force" not in sys.argv:
        print("up to date:", LIB)
        return
    work = "/tmp/magma_ref_build"
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    for dname in ("tools", "include", "control", "magmablas", "src", "interface_cuda"):
        shutil.copytree(os.path.join(REF, dname), os.path.join(work, dname))
    gen = [sys.executable, "tools/codegen.py"]
    hdrs = (glob.glob(work + "/include/*.h") + glob.glob(work + "/control/*.h") + glob.glob(work + "/magmablas/z*.cuh") +
            glob.glob(work + "/magmablas/z*.h") + [work + "/magmablas/commonblas_z.h"])
    hdrs = [os.path.relpath(h, work) for h in hdrs if os.path.exists(h)]
    for i in range(0, len(hdrs), 40):
        sh(gen + hdrs[i:i + 40], cwd=work)
    dsrcs = []
    for dname, files in SRC_D.items():
        for f in files:
            f = f if "." in f else f + ".cpp"
            sh(gen + ["-p", "d", f"{dname}/{f}"], cwd=work)
            dsrcs.append(f"{dname}/{f.replace('z', 'd', 1) if f.startswith('z') else f.replace('_z_', '_d_')}")
    for dname, files in SRC_ALLPREC.items():
        for f in files:
            sh(gen + [f"{dname}/{f}"], cwd=work)
            base = f[1:]
            dsrcs += [f"{dname}/z{base}", f"{dname}/c{base}", f"{dname}/d{base.replace('unm2r', 'orm2r')}",
                      f"{dname}/s{base.replace('unm2r', 'orm2r')}"]
    for dname, files in SRC_PLAIN.items():
        dsrcs += [f"{dname}/{f}" for f in files]
    with open(work + "/include/magma_config.h", "w") as f:
        f.write("#ifndef MAGMA_CONFIG_H\n#define MAGMA_CONFIG_H\n#ifndef MAGMA_HAVE_CUDA\n#define MAGMA_HAVE_CUDA\n#endif\n#endif\n")
    missing = [s for s in dsrcs if not os.path.exists(os.path.join(work, s))]
    if missing:
        raise RuntimeError(f"generated sources missing: {missing}")
    flags = ["-std=c++17", "-O3", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100,code=sm_100",
             "-DMAGMA_HAVE_CUDA", "-DADD_", "-DNDEBUG", "-DMAGMA_CUDA_ARCH_MIN=1000", '-DMAGMA_CUDA_ARCH="sm_100"', "-Iinclude", "-Icontrol", "-Imagmablas", "-x", "cu", "-c"]

    def cc(s):
        o = os.path.join(work, "obj_" + s.replace("/", "_") + ".o")
        sh(["nvcc", *flags, s, "-o", o], cwd=work)
        return o

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, dsrcs))
    os.makedirs(OUT, exist_ok=True)
    sh(["nvcc", "-shared", "-o", LIB, *objs, "-lcublas", "-lcusparse"], cwd=work)
    shutil.rmtree(work, ignore_errors=True)
    print("built", LIB, os.path.getsize(LIB) // 1024, "KiB from", len(dsrcs), "sources")


if __name__ == "__main__":
    main()
