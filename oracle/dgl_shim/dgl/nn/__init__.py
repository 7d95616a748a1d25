"""TEST INFRASTRUCTURE ONLY -- dgl shim namespace (see dgl/__init__.py)."""
from . import pytorch  # noqa: F401
