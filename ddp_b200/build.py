"""Build recipe of libddp_b200.so (sm_100a only, in-tree so that it travels with the repo snapshot)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "ddp_b200.cu")
LIB = os.path.join(PKG, "libddp_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libddp_b200.so cannot be built")
    return p


STAMP = os.path.join(PKG, "libddp_b200.stamp")


def source_digest():
    """sha256 over the CUDA sources + the C header + the flags (mtimes do not survive the snapshot to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    csrc = os.path.join(PKG, "csrc")
    for f in sorted(os.listdir(csrc)) + [os.path.join("..", "..", "include", "ddp_b200.h")]:
        with open(os.path.join(csrc, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != source_digest()


def build(force=False, verbose=False):
    """Compile ddp_b200/csrc/*.cu -> ddp_b200/libddp_b200.so with nvcc for sm_100a."""
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_digest() + "\n")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
