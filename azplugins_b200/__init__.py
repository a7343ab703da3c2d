"""azplugins_b200 -- B200-native neighbour-list pair-force path of mphowardlab/azplugins.

Host-side mirror of the reference plugin surface (``pair``, ``external``, ``wall``), the HOOMD-layout neighbour list
(``nlist``), device particle data (``State``, ``Box``), the multi-GPU particle-slice scheduler
(``slices``) and synthetic workload generators (``synth``). Everything numeric runs in
``libazp_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/azp_b200.h``).
"""

from . import _lib  # noqa: F401  (fails loudly when the CUDA extension is not built)
from . import external, kernels, md, nlist, pair, wall
from .box import Box
from .state import State

__all__ = ["Box", "State", "external", "kernels", "md", "nlist", "pair", "wall"]
