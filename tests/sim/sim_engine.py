"""The product's ctypes wrapper (porechop_b200/cpp_function_wrappers.py) bound to the HOST-SIMULATED engine
(tests/sim/_build/libengine_sim.so, see pbsim_cuda.h / build_sim.py) instead of cpp_functions.so.  TESTS ONLY."""
import ctypes
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_sim  # noqa: E402

_cache = {}


def load():
    if 'w' in _cache:
        return _cache['w']
    from porechop_b200 import cpp_function_wrappers as real
    so = build_sim.build()
    spec = importlib.util.spec_from_file_location('pb200_sim_wrappers', real.__file__)
    mod = importlib.util.module_from_spec(spec)
    orig = ctypes.CDLL

    def cdll(path, *a, **k):
        return orig(so if os.path.basename(str(path)) == 'cpp_functions.so' else path, *a, **k)
    ctypes.CDLL = cdll
    try:
        spec.loader.exec_module(mod)
    finally:
        ctypes.CDLL = orig
    _cache['w'] = mod
    return mod
