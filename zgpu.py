"""Import shim: the product package directory is `spicedb-kubeapi-proxy_b200/` (the name
the project mandates; '-' is not importable), so it is loaded here under the module
name `spicedb_kubeapi_proxy_b200` and re-exported. Usage: `import zgpu`."""
import importlib.util
import os
import sys

_NAME = "spicedb_kubeapi_proxy_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spicedb-kubeapi-proxy_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
from spicedb_kubeapi_proxy_b200 import *  # noqa: F401,F403,E402
from spicedb_kubeapi_proxy_b200 import _lib, client  # noqa: F401,E402
