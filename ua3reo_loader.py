"""Imports the hyphen-named package directory `ua3reo-ddc-transceiver_b200/` as a module."""
import importlib.util
import os
import sys

_NAME = "ua3reo_ddc_transceiver_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(here, "ua3reo-ddc-transceiver_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
