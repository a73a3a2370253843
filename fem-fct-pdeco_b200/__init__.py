"""fem-fct-pdeco_b200: B200-native (sm_100a) implementation of the hot path of
KarolinaBenkova/FEM-FCT-PDECO -- the P1 FEM flux-corrected-transport step, the element assembly feeding it
and the sparse state/adjoint solves of the PDECO loops -- behind the reference's own ``helpers.py`` names.

    import fem_fct_pdeco_b200 as fp          # importable alias of this directory (see fem_fct_pdeco_b200.py)
    from fem_fct_pdeco_b200.helpers import FCT_alg_ref, ChebSI, ...

Importing the package loads ``libfctpdeco.so`` and fails loudly if it has not been built.
"""
from . import _lib                      # noqa: F401  (raises ImportError if the CUDA library is missing)
from ._lib import FctError, device_count
from .context import DeviceArray, FctContext
from .mesh import FunctionSpaceP1, RectMeshP1, vertex_to_dof_map
from . import helpers

__all__ = ["FctError", "device_count", "DeviceArray", "FctContext", "FunctionSpaceP1", "RectMeshP1",
           "vertex_to_dof_map", "helpers"]
