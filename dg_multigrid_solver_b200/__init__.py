"""dg_multigrid_solver_b200 -- B200-native (sm_100a) DG assembly + multigrid V-cycle, a drop-in for
the hot path of thmsdelange/dg-multigrid-solver (`dgfem`).

Host code mirrors the reference's interface (Settings, Geometry/Grid/CoarseGrid, DiscreteSystem,
Relaxation, Solver, DGFEM, `python -m ... -m/-s`); all arithmetic runs in hand-written CUDA
behind the C ABI of include/dgb200.h (libdgb200.so, loaded with ctypes).  PyTorch owns device
buffers and provides torch.distributed; it is plumbing, not the product.  There is no CPU path.
"""
from .settings import Settings, load_params, update_parameters   # noqa: F401

__all__ = ["Settings", "load_params", "update_parameters", "DGFEM", "Solver", "Relaxation",
           "DiscreteSystem", "Geometry", "Grid", "CoarseGrid"]


def __getattr__(name):
    # lazy: importing the package must not require torch/CUDA (the CPU test tier imports it)
    if name == "DGFEM":
        from .dgfem import DGFEM
        return DGFEM
    if name == "Solver":
        from .solver import Solver
        return Solver
    if name == "Relaxation":
        from .relaxation import Relaxation
        return Relaxation
    if name == "DiscreteSystem":
        from .discrete_system import DiscreteSystem
        return DiscreteSystem
    if name in ("Geometry", "Grid", "CoarseGrid"):
        from . import grid
        return getattr(grid, name)
    raise AttributeError(name)
