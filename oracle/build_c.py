"""Build oracle/pencil_c.c -> oracle/libpencil_c.so (gcc, OpenMP) -- TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pencil_c.c")
LIB = os.path.join(HERE, "libpencil_c.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
                        "-o", LIB, SRC], check=True)
    return LIB


if __name__ == "__main__":
    print(build(True))
