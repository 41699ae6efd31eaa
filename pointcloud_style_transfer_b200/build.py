"""Build csrc/libpcst.so in-tree with nvcc for sm_100a (``python -m pointcloud_style_transfer_b200.build``)."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def build(force: bool = False, verbose: bool = False) -> str:
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    lib = os.path.join(CSRC, "libpcst.so")
    if not os.path.exists(lib):
        raise RuntimeError("libpcst.so was not produced")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
