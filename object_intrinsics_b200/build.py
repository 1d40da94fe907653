"""Builds liboi_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m object_intrinsics_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liboi_b200.so")
SOURCES = ["oi_api.cu", "oi_render_ffma.cu", "oi_render_tc.cu", "oi_render_bwd.cu", "oi_wgrad_tc.cu", "oi_render_bwd_tc.cu",
           "oi_render_aux.cu", "oi_ops.cu",
           "oi_generator.cu", "oi_augment.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# NOTE: --use_fast_math is deliberately NOT wanted for the render path (expf/div accuracy); see below.
NVCC_FLAGS.remove("--use_fast_math")


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stamp():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [os.path.join(ROOT, "include", "oi_b200.h")]
    # content and repo-relative names only: the same sources give the same stamp wherever the tree is checked out
    # (profiles/traffic.json refers to it, and the GPU box runs the tree from a scratch path)
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.relpath(f, ROOT).encode() + b"\0" + fh.read())
    h.update(" ".join(a.replace(ROOT, ".") for a in NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, variant=None, defines=()):
    """Default: object_intrinsics_b200/lib/liboi_b200.so.  `variant` (developer tooling for A/B timing on the GPU
    box): the same sources with extra -D `defines`, written to lib/variants/<variant>/liboi_b200.so and selected at
    run time with OI_LIB_PATH (see _lib.py)."""
    lib_dir = LIB_DIR if variant is None else os.path.join(LIB_DIR, "variants", variant)
    lib_path = os.path.join(lib_dir, "liboi_b200.so")
    os.makedirs(lib_dir, exist_ok=True)
    stamp_file = os.path.join(lib_dir, "liboi_b200.stamp")
    stamp = _stamp() + "".join(" -D" + d for d in defines)
    if not force and os.path.exists(lib_path) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return lib_path
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(lib_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *["-D" + d for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    cmd = [_nvcc(), "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    with open(os.path.join(lib_dir, "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if verbose:
        print("\n".join(log))
    return lib_path


if __name__ == "__main__":
    # python -m object_intrinsics_b200.build [--force] [--variant NAME -DFOO=1 -DBAR=2 ...]
    args = sys.argv[1:]
    var = args[args.index("--variant") + 1] if "--variant" in args else None
    print(build(force="--force" in args, verbose=var is None, variant=var,
                defines=[a[2:] for a in args if a.startswith("-D")]))
