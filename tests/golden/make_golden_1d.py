"""Expected matrices / vectors of forms over ONE knot vector (refvform_cases.cases1d): assembled by the REAL
reference through its JIT-compiled assemblers.  Run here (oracle/_ref installed):  python tests/golden/make_golden_1d.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
sys.path.insert(0, HERE)

from pyiga import assemble  # noqa: E402
import refvform_cases as rc  # noqa: E402

out = {}
for name, (problem, kvs, geo, inputs) in rc.cases1d().items():
    A = assemble.assemble(problem if isinstance(problem, str) else problem(), kvs, geo=geo, **inputs)
    out['f1_' + name] = A.toarray() if hasattr(A, 'toarray') else np.asarray(A)
np.savez_compressed(os.path.join(HERE, 'ref_forms_1d.npz'), **out)
print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
