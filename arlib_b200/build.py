"""Build libagcf.so in-tree (nvcc, sm_100a only):  python -m arlib_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=True, force=False):
    csrc = os.path.join(HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, stdout=subprocess.DEVNULL)
    res = subprocess.run(["make", "-C", csrc, "-j4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libagcf.so failed")
    return os.path.join(HERE, "libagcf.so")


if __name__ == "__main__":
    build(force="--force" in sys.argv)
