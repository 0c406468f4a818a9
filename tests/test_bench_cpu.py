"""bench.py's reference arm (the real reference on the host cores) at a tiny size: the JSON line keeps the
contract of the driver (impl, metric, unit, e2e, cpu_baseline with the size that was timed)."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT


def test_reference_arm_line():
    if not os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'pyiga')):
        pytest.skip('oracle/_ref is not installed')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--ref-n', '8', '--ref-quick',
                        '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'nnz/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['n_gpus'] == 1 and d['dtype'] == 'f64'
    assert 'n=8' in d['config']['workload']                    # the size that was timed, not the GPU arm's
    cb = d['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['ref_n'] == 8 and abs(cb['value'] - d['value']) <= 1e-9 * d['value']
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0 and abs(e['value'] - d['value']) <= 1e-9 * d['value']
