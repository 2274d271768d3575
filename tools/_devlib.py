"""Development tools load the experiments build of the library (`make -C scpn_quantum_control_b200/csrc experiments`,
tools/_build/libxyk_exp.so: scheduler environment overrides, ablations, per-pass print-outs) or the file named by
XYK_LIB_AB; the shipped scpn_quantum_control_b200/libxyk.so reads no environment variables."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def use_dev_library(default: str = "libxyk_exp.so") -> str:
    from scpn_quantum_control_b200 import _lib

    path = os.environ.get("XYK_LIB_AB") or os.path.join(ROOT, "tools", "_build", default)
    if os.path.exists(path):
        _lib.LIB_PATH = os.path.abspath(path)
    return _lib.LIB_PATH
