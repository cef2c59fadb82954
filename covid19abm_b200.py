"""Import shim: the package directory is named `covid19abm.jl_b200` (with a dot), which Python's import
statement cannot spell. `import covid19abm_b200 as cv` loads it from that directory."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "covid19abm.jl_b200")
_spec = importlib.util.spec_from_file_location("covid19abm_jl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["covid19abm_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})

import importlib as _il
local_run = _il.import_module("covid19abm_jl_b200.local_run")     # scripts/local_run.jl mirror (run, calibrate, ...)
