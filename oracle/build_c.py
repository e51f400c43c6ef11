"""TEST INFRASTRUCTURE -- builds the C restatement of the oracle (oracle/heli3dof_ref.c) with gcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "heli3dof_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libheli3dof_ref.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -ffp-contract=off: no FMA contraction, the literal operator sequence of the reference
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


SRC6 = os.path.join(HERE, "heli6dof_ref.c")
LIB6 = os.path.join(OUT_DIR, "libheli6dof_ref.so")


def build6(force=False):
    """The 6-DOF restatement (oracle/heli6dof_ref.c)."""
    if not force and os.path.exists(LIB6) and os.path.getmtime(LIB6) >= os.path.getmtime(SRC6):
        return LIB6
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB6, SRC6, "-lm"]
    subprocess.check_call(cmd)
    return LIB6


if __name__ == "__main__":
    print(build(force=True))
    print(build6(force=True))
