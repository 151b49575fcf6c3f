"""`oriana` -- import-path alias so that code written against the reference package
(`from oriana.models import ZIGaP`, `from oriana.nodes import Gamma`, `from oriana import Dimensions`, ...)
runs on the B200 implementation unchanged.  Everything lives in `oriana_b200`."""
import importlib
import sys

import oriana_b200 as _impl
from oriana_b200 import *  # noqa: F401,F403

for _name in ('exceptions', 'parameters', 'dims', 'utils', 'nodes', 'models', 'inference', 'singlecell'):
    _mod = importlib.import_module('oriana_b200.' + _name)
    sys.modules[__name__ + '.' + _name] = _mod
    globals()[_name] = _mod
__version__ = _impl.__version__
