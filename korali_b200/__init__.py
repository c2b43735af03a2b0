"""korali_b200 — a B200-native CMA-ES generation loop behind Korali's own API.

    import korali_b200 as korali
    e = korali.Experiment(); e["Solver"]["Type"] = "Optimizer/CMAES"; ...; korali.Engine().run(e)

Only the path `"Solver": {"Type": "Optimizer/CMAES"}` of `"Problem": {"Type": "Optimization"}` is served (SURVEY.md scope).
Layers: korali_b200._host (pybind11, C++: Engine / Experiment, mirrors python/korali/__init__.py:9-20 + source/engine.cpp:201-254)
 -> include/kcma.h (C ABI) -> korali_b200/libkcma.so (hand-written sm_100a CUDA). There is no CPU fallback.
"""
import os as _os

_HERE = _os.path.dirname(_os.path.abspath(__file__))


def _load_host():
    import importlib
    from . import _lib
    _lib.lib()   # fail loudly when libkcma.so is missing
    try:
        return importlib.import_module("korali_b200._host")
    except ImportError as exc:   # pragma: no cover
        raise ImportError("korali_b200/_host*.so is missing: run `python -m korali_b200.build`") from exc


def Engine():
    """korali.Engine() (python/korali/__init__.py:9-11)."""
    return _load_host().Engine()


def Experiment():
    """korali.Experiment() (python/korali/__init__.py:14-16)."""
    return _load_host().Experiment()
