"""Expected matrices for tests of the reference-VForm backend (pyiga_b200/refvform.py): the REAL reference
assembles the forms of refvform_cases.py with its JIT-compiled Cython assemblers, and its HDiscretization
assembles THB / HB stiffness matrices.  Run here (oracle/_ref installed):  python tests/golden/make_golden_vform.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
sys.path.insert(0, HERE)

from pyiga import _hdiscr, assemble, geometry  # noqa: E402
import refvform_cases as rc  # noqa: E402

out = {}
for name, (make, kvs, geo, inputs) in rc.cases().items():
    A = assemble.assemble(make(), kvs, geo=geo, **inputs)
    out['vf_' + name] = A.toarray() if hasattr(A, 'toarray') else np.asarray(A)
for name, (make, kvs, geo, inputs, sides) in rc.bcases().items():
    for bd in sides:
        A = assemble.assemble(make(), kvs, geo=geo, boundary=bd, **inputs)
        out['bd_%s_%s' % (name, bd)] = A.toarray() if hasattr(A, 'toarray') else np.asarray(A)
for name, (make, kvs2, geo, inputs) in rc.pgcases().items():
    out['pg_' + name] = assemble.assemble(make(), kvs2, geo=geo, **inputs).toarray()
geo = geometry.bspline_quarter_annulus()
for name, (make, inputs, sym) in rc.hcases().items():
    for truncate in (False, True):
        hs = rc.hspace(truncate)
        out['h_%s_%d' % (name, truncate)] = _hdiscr.HDiscretization(hs, make(), dict(inputs, geo=geo)).assemble_matrix(symmetric=sym).toarray()
np.savez_compressed(os.path.join(HERE, 'ref_vform_objects.npz'), **out)
print({k: v.shape for k, v in out.items()})
