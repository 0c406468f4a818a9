"""Run the REFERENCE's own test functions against pyiga_b200: the names `pyiga.<module>` are bound to the modules
of this package, the test file is executed unmodified (optionally with textual substitutions given as
--sub OLD=NEW, e.g. to pick spline degrees the device tables cover), and every `test_*` function is called.

    python tools/run_reference_tests.py [--backend emu|cuda] [--sub OLD=NEW ...] /root/reference/test/test_assemble.py [names ...]

Prints one JSON object {test name: "ok" | "FAIL line N: Error: message"}.  `--backend emu` routes the package
through the sequential host emulation of its kernels (tests/emu: CPU-only containers); `cuda` is the product.
The reference test files are read where they lie; nothing of them is copied."""
import argparse
import importlib
import json
import os
import sys
import traceback
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

MODULES = ('bspline', 'geometry', 'assemble', 'vform', 'utils', 'approx', 'operators', 'mlmatrix', 'quadrature', 'assemblers')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--backend', default='emu', choices=['emu', 'cuda'])
    ap.add_argument('--mode', default='alias', choices=['alias', 'patch'],
                    help="alias: pyiga.* are this package's modules; patch: the real reference with its assembly entry points rebound")
    ap.add_argument('--sub', action='append', default=[])
    ap.add_argument('--skip', action='append', default=[], help='test function not to run')
    ap.add_argument('testfile')
    ap.add_argument('names', nargs='*')
    a = ap.parse_args()
    import numpy as np
    import scipy.sparse
    from pyiga_b200 import _device
    if a.backend == 'emu':
        from emu import build_emu
        from emu.emu_backend import EmuBackend
        _device._backend = EmuBackend(build_emu.build())
    if a.mode == 'patch':
        # the REAL reference (oracle/_ref) with its assembly entry points rebound to this package: what a maintainer
        # who integrates the library would do; the reference's other modules (hierarchical spaces,
        # solvers, ...) then drive the device assemblers
        sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
        import pyiga
        import pyiga.assemble
        import pyiga.assemblers
        import pyiga.compile
        from pyiga_b200 import assemble as oa, assemblers as oas, vform as ovf
        sys.modules['pyiga.assemblers'] = oas
        pyiga.assemblers = oas
        pyiga.assemble.assemblers = oas
        pyiga.compile.compile_vform = ovf.compile_vform
        # (pyiga.assemble.assemble itself stays: it dispatches hierarchical spaces and otherwise calls the two below)
        for name in ('mass', 'stiffness', 'divdiv', 'assemble_entries', 'assemble_entries_vec',
                     'instantiate_assembler', 'Assembler', 'inner_products', 'bsp_mass_2d', 'bsp_mass_3d', 'bsp_stiffness_2d',
                     'bsp_stiffness_3d', 'integrate'):
            setattr(pyiga.assemble, name, getattr(oa, name))
    else:
        shim = types.ModuleType('pyiga')
        shim.__path__ = []
        sys.modules['pyiga'] = shim
        for name in MODULES:
            mod = importlib.import_module('pyiga_b200.' + name)
            sys.modules['pyiga.' + name] = mod
            setattr(shim, name, mod)

    def read_sparse_matrix(fname):      # loader of the reference's golden .mtx.gz files (pyiga/utils.py:54-60): test I/O only
        I, J, vals = np.loadtxt(fname, skiprows=1, unpack=True)
        return scipy.sparse.coo_matrix((vals, (I.astype(int) - 1, J.astype(int) - 1))).tocsr()
    if a.mode == 'alias' and not hasattr(sys.modules['pyiga.utils'], 'read_sparse_matrix'):
        sys.modules['pyiga.utils'].read_sparse_matrix = read_sparse_matrix
    src = open(a.testfile).read()
    for sub in a.sub:
        old, new = sub.split('=', 1)
        assert old in src, 'substitution %r does not apply' % old
        src = src.replace(old, new)
    g = {'__name__': 'reference_test', '__file__': a.testfile}
    exec(compile(src, a.testfile, 'exec'), g)
    res = {}
    for k, fn in list(g.items()):
        if k.startswith('test_') and callable(fn) and (not a.names or k in a.names) and k not in a.skip:
            try:
                fn()
                res[k] = 'ok'
            except BaseException as e:      # noqa: BLE001 - report, do not stop
                tb = [t for t in traceback.extract_tb(e.__traceback__) if t.filename == a.testfile]
                res[k] = 'FAIL line %s: %s: %s' % (tb[-1].lineno if tb else '?', type(e).__name__, str(e)[:160])
    print(json.dumps(res))


if __name__ == '__main__':
    main()
