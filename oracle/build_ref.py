"""Builds the REAL reference (c-f-h/pyiga) from /root/reference into oracle/_ref/ (git-ignored).

TEST / BASELINE INFRASTRUCTURE.  The install is the checker that pins the oracle (it generated
tests/golden/ref_cases.npz) and the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).  It is
built from an unmodified scratch copy of the sources with the reference's own setup.py and flags
(-O3 -march=native -ffast-math -fopenmp, setup.py:11-17); CC is overridden because the image's
default gcc wrapper cannot link -fopenmp.  Nothing under pyiga_b200/ imports it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SRC = '/root/reference'


def available():
    return os.path.isdir(os.path.join(DEST, 'pyiga'))


def build(force=False):
    if available() and not force:
        return DEST
    if not os.path.isdir(SRC):
        raise RuntimeError('reference sources not present at %s' % SRC)
    tmp = tempfile.mkdtemp(prefix='pyiga_src_')
    try:
        work = os.path.join(tmp, 'src')
        shutil.copytree(SRC, work)
        subprocess.run(['chmod', '-R', 'u+w', work], check=True)
        env = dict(os.environ, CC='/usr/bin/gcc', CXX='/usr/bin/g++', LDSHARED='/usr/bin/gcc -shared')
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        subprocess.run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
                        '--find-links', '/opt/wheelhouse', '--target', DEST, work], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return DEST


def import_reference():
    """Import the installed reference package (raises ImportError if it is not there)."""
    if not available():
        raise ImportError('oracle/_ref is not built')
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import pyiga
    return pyiga


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
