"""The mesh readers on damaged files (host_tau.cpp, host_foam.cpp): truncations, damaged header bytes, damaged counts and
labels.  A reader answers with a mesh or with an error message -- never a crash, and never an allocation sized by a damaged
count (the address space of the child process is capped at 3 GB; an OpenFOAM list size, a patch's nFaces or a cell label
larger than the file allows is refused before it sizes anything)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind,seed,count", [("tau", 7, 240), ("foam", 7, 240)])
def test_damaged_files_end_in_a_mesh_or_a_message(tmp_path, kind, seed, count):
    out = subprocess.run([sys.executable, "-m", "tests.reader_fuzz", kind, str(seed), str(count), str(tmp_path)], cwd=ROOT,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    assert "%s: %d mutations" % (kind, count) in out.stdout and " 0 reported as errors" not in out.stdout, out.stdout[-500:]
