"""Host-emulation build of the item kernels + host runtime of the Co-Mod-GAN / conv2d_resample path.

TEST INFRASTRUCTURE ONLY.  ``mi-gan_b200/csrc/comodgan_abi.cu`` is compiled as plain C++ with ``-DMIGAN_EMULATE``:
``ck_launch`` becomes a host loop over the very same kernel functors (``comod_kernels.cuh``) and the GEMM a reference
triple loop, so the index math, padding bookkeeping, weight packing and the whole host walk can be checked against the
oracle on a machine without a GPU (``tests/test_comodgan_emul.py``).  The package never loads this library and the
product library contains no host loop (it is compiled without the macro).
"""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "mi-gan_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcomod_emul.so")
SRCS = ["comodgan_abi.cu", "prepost.cu", "pipeline.cu", "reparam.cu", "comod_kernels.cuh", os.path.join("..", "..", "include", "comodgan_b200.h")]


def _hash() -> str:
    h = hashlib.sha256()
    for s in SRCS:
        with open(os.path.join(CSRC, s), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    stamp = os.path.join(OUT, "stamp")
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _hash():
        return LIB
    cmd = ["g++", "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-DMIGAN_EMULATE", "-x", "c++",
           "comodgan_abi.cu", "prepost.cu", "pipeline.cu", "reparam.cu", "-o", LIB]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + proc.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(_hash())
    return LIB


if __name__ == "__main__":
    print(build())
