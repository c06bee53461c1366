"""Run unmodified user code of the reference's Python package on this library:

    import bess_b200.compat; bess_b200.compat.install_as_bess()
    from bess.linear import PdasLm, L0L2Logistic, GroupPdasCox      # python/bess/linear.py:434-925
    from bess.cbess import pywrap_bess                               # the SWIG entry, python/bess/cbess.py:65-66

``install_as_bess`` registers this package's modules under the reference's module names (``bess``, ``bess.linear``,
``bess.cbess``, ``bess.gen_data``) in ``sys.modules``.  ``bess.gen_data`` is a module with the REFERENCE's signature
(``gen_data(n, p, family, k, rho=0, sigma=1, beta=None, censoring=True, c=1, scal=10)`` and its ``data`` record; cox ``y``
is the unsorted [time, status] array) -- ``bess_b200.gen_data.gen_data`` has other arguments (seed, snr) and is not what
reference user code expects.  It refuses to shadow an already imported real ``bess``."""
from __future__ import annotations

import sys
import types


def install_as_bess(force: bool = False):
    from . import cbess, gen_data, linear
    if "bess" in sys.modules and not force and not getattr(sys.modules["bess"], "__bess_b200__", False):
        raise ImportError("a package named 'bess' is already imported; pass force=True to replace it in this process")
    pkg = types.ModuleType("bess")
    pkg.__doc__ = "bess (Mamba413/bess Python API) served by bess_b200 on sm_100a"
    pkg.__path__ = []  # a package, so that `import bess.linear` resolves through sys.modules
    pkg.__bess_b200__ = True
    gd = types.ModuleType("bess.gen_data")
    gd.__doc__ = "python/bess/gen_data.py of the reference, served by bess_b200.gen_data.gen_data_reference"
    gd.gen_data, gd.data, gd.np = gen_data.gen_data_reference, gen_data.data, gen_data.np
    pkg.linear, pkg.cbess, pkg.gen_data = linear, cbess, gd
    sys.modules["bess"] = pkg
    sys.modules["bess.linear"] = linear
    sys.modules["bess.cbess"] = cbess
    sys.modules["bess.gen_data"] = gd
    return pkg
