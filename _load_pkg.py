"""Loads the package directory `point-cloud-preprocessing-tools_b200/` (not a valid Python identifier) under the
alias `pcpt_b200`, and puts oracle/ (test infrastructure) on the path for the callers that are allowed to use it."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "point-cloud-preprocessing-tools_b200")


def load_pkg():
    if "pcpt_b200" in sys.modules:
        return sys.modules["pcpt_b200"]
    spec = importlib.util.spec_from_file_location("pcpt_b200", os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pcpt_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_synth():
    load_pkg()
    import importlib
    return importlib.import_module("pcpt_b200.synth")


def load_oracle():
    """TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs call this."""
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import oracle_lib
    return oracle_lib
