"""Import alias: ``import b200ddp`` loads the package that lives in ``./distributed-pytorch_b200/``.

The package directory name is fixed by the build contract and contains a hyphen, which the ``import``
statement cannot spell.  This shim loads that directory exactly once under the importable name
``b200ddp`` (sub-modules resolve as ``b200ddp.<name>``), so there is a single copy of every module.
"""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "distributed-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "b200ddp", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200ddp"] = _mod
_spec.loader.exec_module(_mod)
