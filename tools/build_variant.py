"""Development builds of the library beside the shipped one: `python tools/build_variant.py NAME [-DFLAG ...]` writes
tools/_build/libvar_NAME.so (git-ignored; travels to the GPU box).  Select it with SLS_B200_LIB=tools/_build/libvar_NAME.so --
same-box A/B runs are the only comparison that resolves a few per cent (box-to-box variance is +-7 %)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sls_sat_solving_with_deep_learning_b200 import _lib  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "tools", "_build", f"libvar_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
subprocess.check_call([_lib._nvcc(), *_lib.NVCC_FLAGS, *flags, "-o", out, *_lib._SOURCES], cwd=_lib._PKG)
print(out)
