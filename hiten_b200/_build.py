"""Build the sm_100a shared library (C ABI in include/hiten_b200.h) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhiten_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr",
    "-t", "8",
    "-Xcompiler", "-fPIC",
    "-shared",
    "-ldl",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(HERE, "..", "include", "hiten_b200.h"))
    return out


HASH_PATH = LIB_PATH + ".srchash"


def source_hash():
    """Content hash of everything the library is built from (sources, public header, flags): modification times do
    not survive the copy to a GPU box, contents do."""
    import hashlib
    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    for d in sorted(_deps()):
        if os.path.isfile(d):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was built from other sources than the ones on disk."""
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(HASH_PATH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def have_nvcc():
    import shutil
    return os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) or shutil.which("nvcc") is not None


def build(force=False, verbose=False, extra_flags=(), out=None):
    if out is None and not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    target = LIB_PATH if out is None else out
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + sources()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if out is None and not extra_flags:
        with open(HASH_PATH, "w") as f:
            f.write(source_hash())
    return target


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
