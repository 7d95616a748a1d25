"""Import alias: ``import gnn_recsys_b200`` loads the package directory ``gnn-recsys_b200/``.

The directory name is fixed by the build contract and is not a valid Python identifier, so this module
loads it through importlib and registers the package (and its sub-modules) under the underscore name.
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module('gnn-recsys_b200')
for _name, _mod in list(sys.modules.items()):
    if _name == 'gnn-recsys_b200' or _name.startswith('gnn-recsys_b200.'):
        sys.modules[_name.replace('gnn-recsys_b200', 'gnn_recsys_b200', 1)] = _mod
