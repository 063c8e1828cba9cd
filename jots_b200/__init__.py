"""jots_b200 -- B200-native implicit conduction solve for JOTS (ctypes binding of libjots_b200.so).

The product is the C-ABI shared library built from ``jots_b200/csrc`` (include/jots_b200.h);
this package only loads it and mirrors the reference's operator / driver interface for tests
and benchmarks.  There is no CPU fallback: without the built library, or without a CUDA device,
every compute entry point raises.
"""
from .lib import (JotsError, config_dump, patch_shape, h1_dof_map, gather_from_native, mfem_dof_numbering, restart_write, restart_read, paraview_write, vtk_lagrange_order, Mesh, Space, Context, Operator, Driver, Partition, Comm, load_library, library_path,
                  HOST, DEVICE, UNIFORM, POLYNOMIAL, BC_ISOTHERMAL, BC_HEATFLUX, BC_SIN_ISOTHERMAL,
                  BC_SIN_HEATFLUX, CG, GMRES, FGMRES, JACOBI, CHEBYSHEV, EULER_IMPLICIT, EULER_EXPLICIT,
                  RK4, LINEARIZED_UNSTEADY, NONLINEAR_UNSTEADY, STEADY, DENSITY, SPECIFIC_HEAT,
                  THERMAL_CONDUCTIVITY, FORM_MASS, FORM_STIFFNESS, FORM_SYSTEM, FORM_RESIDUAL)

__all__ = [n for n in dir() if not n.startswith("_")]
