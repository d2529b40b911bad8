"""Device-resident stand-in for ``fenics_constitutive.solver`` (reference
src/fenics_constitutive/solver/__init__.py:1-10) plus the few pieces of dolfinx
it needs (mesh, function space, Dirichlet BCs, Newton solver).  dolfinx/PETSc
are absent from the build image; see DESIGN.md "stand-in driver"."""
from __future__ import annotations

from ._newton import KrylovError, NewtonSolver
from ._problem import (
    History,
    IncrementalDisplacement,
    IncrementalStress,
    IncrSmallStrainProblem,
    LawOnSubMesh,
    QuadratureFunction,
    SimulationTime,
)
from .maps import IdentityMap, SpaceMap, SubSpaceMap, build_subspace_map
from .partitioned import MeshPartition
from .mesh import (
    Constant,
    DirichletBC,
    ElementTables,
    Function,
    FunctionSpace,
    Mesh,
    create_box,
    create_rectangle,
    create_unit_cube,
    create_unit_interval,
    create_unit_square,
    dirichletbc,
    functionspace,
    locate_dofs_geometrical,
    surface_load,
)

__all__ = [
    "IncrSmallStrainProblem", "KrylovError", "NewtonSolver", "SimulationTime", "IncrementalDisplacement",
    "IncrementalStress", "History", "LawOnSubMesh", "QuadratureFunction", "IdentityMap", "SpaceMap", "SubSpaceMap",
    "build_subspace_map", "Mesh", "FunctionSpace", "Function", "Constant", "DirichletBC", "ElementTables",
    "create_unit_interval", "create_unit_square", "create_rectangle", "create_unit_cube", "create_box",
    "functionspace", "dirichletbc", "locate_dofs_geometrical", "MeshPartition", "surface_load",
]
