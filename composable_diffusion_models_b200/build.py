"""Build libcdm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m composable_diffusion_models_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcdm_b200.so")
STAMP = os.path.join(HERE, ".libcdm_b200.stamp")
SOURCES = ["api.cu", "step_kernels.cu", "elementwise.cu", "init_conv_tc.cu", "conv_fp32.cu", "conv_tc.cu", "conv_tc2.cu", "conv_tc3.cu", "conv_x3.cu", "unet.cu", "jvp.cu", "mlp.cu", "mlp_tc.cu", "general_fp32.cu", "experts2.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"] + os.environ.get("CDM_NVCC_EXTRA", "").split()


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _fingerprint():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "cdm_b200.h")]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def abi_stamp():
    """First 32 bits of sha256(include/cdm_b200.h): compiled into the library (cdm_abi_stamp) and checked at load, so a
    signature edit can never meet a .so built from the older header."""
    hdr = os.path.join(HERE, "..", "include", "cdm_b200.h")
    return int(hashlib.sha256(open(hdr, "rb").read()).hexdigest()[:8], 16)


def _up_to_date(fp, force):
    return not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp


def build(force=False, verbose=False):
    fp = _fingerprint()
    if _up_to_date(fp, force):
        return LIB
    # one builder at a time (torchrun starts every rank at once): the others wait, then find the library up to date
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if _up_to_date(fp, force):
                return LIB
            return _build_locked(fp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(fp, verbose):
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + [f"-DCDM_ABI_STAMP={abi_stamp()}u"]
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, "-shared", "-cudart", "static", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
