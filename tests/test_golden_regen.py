"""The committed golden fixtures are exactly what tests/golden/make_golden.py writes TODAY from the reference's own
code: when /root/reference is present (the build container; it does not exist on the GPU box) every generator group is
re-run into a temporary directory and every array of every fixture is compared bit for bit with the committed file.
A fixture that drifted from its script (or a script that changed without regenerating) fails here."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import GOLDEN, ROOT

REFERENCE = os.environ.get('GNN_RECSYS_REFERENCE', '/root/reference')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'src')), reason='reference sources not present on this box')
def test_committed_fixtures_regenerate_bit_exactly(tmp_path):
    out = str(tmp_path / 'golden')
    code = ("import sys; sys.path.insert(0, %r); import make_golden as M; M.generate(%r)" % (GOLDEN, out))
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    fresh = sorted(os.path.basename(p) for p in glob.glob(os.path.join(out, '*.npz')))
    committed = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, '*.npz')))
    assert fresh == committed, set(fresh) ^ set(committed)
    for name in committed:
        a, b = np.load(os.path.join(GOLDEN, name)), np.load(os.path.join(out, name))
        assert sorted(a.files) == sorted(b.files), name
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, (name, k)
            assert np.array_equal(a[k], b[k], equal_nan=True), '%s: array %s differs from what make_golden.py writes' % (name, k)
