"""zk-toolkit_b200: B200-native MSM / Groth16-prove path for zk-toolkit's BLS12-381.

The directory name carries a hyphen (as the project is named); import it with
``importlib.import_module("zk-toolkit_b200")`` or through the ``zk_toolkit_b200`` shim at the
repository root.  All compute happens in ``libzkmsm.so`` (hand-written sm_100a CUDA behind the C
ABI of ``include/zkmsm.h``); there is no CPU fallback.
"""
from ._lib import LIB_PATH, SYMBOLS, ZkmsmError, load  # noqa: F401
from .context import Context, PointSet  # noqa: F401
from .api import (G1Point, G1Points, G2Point, G2Points, Polynomial, Q, R, default_context,  # noqa: F401
                  scalars_to_array)
from . import groth16, pinocchio, synthetic  # noqa: F401,E402
