"""Builds the oracle's C restatement (oracle/_build/liboracle.so) and, when /root/reference is present, the
reference's own sources into oracle/_ref/ (never copied into the repo; outputs are git-ignored but travel with gpurun).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REFDIR = os.path.join(HERE, "_ref")
LIB = os.path.join(BUILD, "liboracle.so")
REFERENCE = "/root/reference"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _newer(dst, srcs):
    return os.path.exists(dst) and all(os.path.getmtime(dst) >= os.path.getmtime(s) for s in srcs)


def build_oracle(force=False):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "csrc", "oracle.c")
    if force or not _newer(LIB, [src]):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", "-o", LIB, src, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stderr)
    return LIB


def _build_ref_lib(out, wrap, ref_srcs, inc_dirs, force, verbose):
    srcs = [wrap] + ref_srcs
    if not force and _newer(out, srcs):
        return out
    # some of the reference .cu files include headers that pull in <torch/serialize/tensor.h>: give nvcc torch's include dirs
    import sysconfig
    import torch.utils.cpp_extension as ext
    inc = []
    for p in ext.include_paths():
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    for d in inc_dirs:
        inc += ["-I", d]
    objs = []
    tag = os.path.basename(out)[3:-3]
    os.makedirs(os.path.join(REFDIR, "obj"), exist_ok=True)
    for i, s in enumerate(srcs):
        o = os.path.join(REFDIR, "obj", "%s_%d_%s.o" % (tag, i, os.path.basename(s)[:-3]))
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-w",
               *inc, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference kernel build failed for %s:\n%s" % (s, r.stderr[-2000:]))
        objs.append(o)
    r = subprocess.run([NVCC, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference kernel link failed:\n" + r.stderr[-2000:])
    if verbose:
        print("built", out)
    return out


def build_ref(force=False, verbose=False):
    """Compiles the reference's CUDA kernels (from where they lie under /root/reference) behind extern "C" wrappers
    (oracle/ref_wrap/*.cu) into oracle/_ref/libpcdet_ref_kernels.so (iou3d_nms, roiaware_pool3d, pointnet2_stack) and
    oracle/_ref/libpcdet_ref_kernels_batch.so (pointnet2_batch, voxel_query, roipoint_pool3d; see build_ref_batch). Returns the
    path of the first or None when the reference tree is absent (GPU box: the prebuilt files travel with the snapshot)."""
    out = os.path.join(REFDIR, "libpcdet_ref_kernels.so")
    if not os.path.isdir(REFERENCE):
        return out if os.path.exists(out) else None
    os.makedirs(REFDIR, exist_ok=True)
    ops = os.path.join(REFERENCE, "pcdet", "ops")
    pn = os.path.join(ops, "pointnet2", "pointnet2_stack", "src")
    srcs = [
        os.path.join(ops, "iou3d_nms", "src", "iou3d_nms_kernel.cu"),
        os.path.join(ops, "roiaware_pool3d", "src", "roiaware_pool3d_kernel.cu"),
    ] + [os.path.join(pn, f) for f in ("ball_query_gpu.cu", "group_points_gpu.cu", "sampling_gpu.cu", "interpolate_gpu.cu")]
    _build_ref_lib(out, os.path.join(HERE, "ref_wrap", "ref_kernels_wrap.cu"), srcs, [pn], force, verbose)
    build_ref_batch(force, verbose)
    return out


def build_ref_batch(force=False, verbose=False):
    """The second reference library: pointnet2_batch (its sampling_gpu.cu defines the same launcher name as the stack one,
    hence a library of its own), voxel_query_gpu.cu of pointnet2_stack and roipoint_pool3d_kernel.cu."""
    out = os.path.join(REFDIR, "libpcdet_ref_kernels_batch.so")
    if not os.path.isdir(REFERENCE):
        return out if os.path.exists(out) else None
    os.makedirs(REFDIR, exist_ok=True)
    ops = os.path.join(REFERENCE, "pcdet", "ops")
    pb = os.path.join(ops, "pointnet2", "pointnet2_batch", "src")
    srcs = [os.path.join(pb, f) for f in ("ball_query_gpu.cu", "group_points_gpu.cu", "sampling_gpu.cu", "interpolate_gpu.cu")]
    ps = os.path.join(ops, "pointnet2", "pointnet2_stack", "src")
    srcs += [os.path.join(ps, "voxel_query_gpu.cu"), os.path.join(ps, "vector_pool_gpu.cu"),
             os.path.join(ops, "roipoint_pool3d", "src", "roipoint_pool3d_kernel.cu")]
    return _build_ref_lib(out, os.path.join(HERE, "ref_wrap", "ref_kernels_batch_wrap.cu"), srcs, [pb, ps], force, verbose)


STAGE = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def stage_reference_python(force=False):
    """Stages the reference's pure-Python package (pcdet/**/*.py) and its model configs (tools/cfgs/**/*.yaml) under
    git-ignored baseline/_ref/ - verbatim copies, never part of the repo's history - so that the GPU box (which has no
    /root/reference) can run the reference's UNMODIFIED wrappers, modules and detector classes over the drop-in
    (tests/ref_env.py). Returns the staged root or None when neither the reference nor a previous staging exists."""
    import shutil
    marker = os.path.join(STAGE, "pcdet", "models", "detectors", "detector3d_template.py")
    if not os.path.isdir(REFERENCE):
        return STAGE if os.path.exists(marker) else None
    if os.path.exists(marker) and not force:
        return STAGE
    for sub, pat in (("pcdet", ".py"), (os.path.join("tools", "cfgs"), ".yaml")):
        for root, _, files in os.walk(os.path.join(REFERENCE, sub)):
            for f in files:
                if f.endswith(pat):
                    src = os.path.join(root, f)
                    dst = os.path.join(STAGE, os.path.relpath(src, REFERENCE))
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(src, dst)
    return STAGE


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv, verbose=True))
    print(stage_reference_python(force="--force" in sys.argv))
