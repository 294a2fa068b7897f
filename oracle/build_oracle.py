"""TEST INFRASTRUCTURE ONLY — compile oracle/philox_oracle.c into oracle/_build/libphilox_oracle.so.

The reference is pure Python (no C sources to compile); its unmodified package is installed into
oracle/_ref by oracle/build_ref.py (timed by bench.py, compared with the restatement by
tests/test_reference_install.py) and exercised at golden-generation time (tests/golden/make_goldens.py).
"""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "philox_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libphilox_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(OUT) or os.path.getmtime(SRC) > os.path.getmtime(OUT):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", OUT, SRC, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
