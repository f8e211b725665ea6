"""Build the plain-C oracle (oracle/rd3_oracle.c -> oracle/_build/librd3_oracle.so).

TEST INFRASTRUCTURE ONLY (see the header of rd3_oracle.c).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rd3_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "librd3_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-mfma", "-ffp-contract=off",
           "-Wall", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
