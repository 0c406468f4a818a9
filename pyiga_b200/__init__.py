"""pyiga_b200 — B200-native tensor-product Gauss-quadrature assembly behind the pyiga API.

One hot path of c-f-h/pyiga, rebuilt for sm_100a: ``assemble.mass/stiffness``, ``assemble.assemble``
and the assembler ``multi_entries`` protocol, returning scipy CSR and multi-level banded layouts.
See DESIGN.md.
"""
__version__ = '0.1.0'

from . import bspline, geometry, quadrature, mlmatrix, assemblers, assemble, operators, utils, approx  # noqa: F401
