"""In-tree builds: the product library (nvcc, sm_100a) and the CPU oracle (g++).  Used by
__graft_entry__.build(); cross-compiles without a GPU."""
import os
import subprocess

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def build_product(jobs: int = 4) -> str:
    _run(["make", "-j%d" % jobs], os.path.join(_REPO, "deepmimic_b200", "csrc"))
    return os.path.join(_REPO, "deepmimic_b200", "libdeepmimic_b200.so")


def build_oracle() -> str:
    _run(["make"], os.path.join(_REPO, "oracle"))
    return os.path.join(_REPO, "oracle", "libdm_oracle.so")
