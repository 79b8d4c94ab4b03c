"""The command-line driver with the particle tracker switched on (csrc/fjsph_run.cpp: FJSPH.cpp:206-210,319-320 and
Integration.cpp:151-169 on the C ABI)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests.test_gpu_driver import BIN, ROOT

pytestmark = pytest.mark.gpu


def test_driver_tracks_the_deleted_particles(tmp_path):
    """FJSPH.cpp:206-210,319-320 + Integration.cpp:151-169 through the driver: the jet deck with a delete plane 5 dx above the
    aero entry plane, coupled to a TAU mesh with a sheared cross flow, `Transition to IPT (0/1): 1`.  Every particle that passes
    the plane is erased from the SPH set, followed through the mesh by the device tracker (csrc/ipt.cu) and written to
    <prefix>_IPT_streaks.dat as one Tecplot zone (ASCII::Write_Streaks, IPT.cpp:205-224)."""
    import re

    from tests.tau_case import write_tau

    for f in ("jet3d.para", "jet3d_fluid.bmap", "jet3d_pipe.bmap"):
        shutil.copy(os.path.join(ROOT, "tests", "decks", f), tmp_path / f)
    fluid = tmp_path / "jet3d_fluid.bmap"
    fluid.write_text(fluid.read_text().replace(" block end", "            Deletion normal: 0,1,0\n    Deletion plane constant: 0.0006\n block end"))
    lo, hi = np.array([-0.004, -0.0012, -0.003]), np.array([0.012, 0.006, 0.003])
    mesh, sol, *_ = write_tau(tmp_path, lo, hi, (16, 9, 6), lambda x: (128.5 + 4000.0 * x[1], 0.0, 0.0), lambda x: 101325.0,
                              lambda x: 1.225, wall_marker=-2)
    (tmp_path / "tau.bmap").write_text(" block begin\n   Markers: 1\n   Type: farfield\n block end\n")
    para = tmp_path / "jet3d.para"
    para.write_text(para.read_text() + "\n SPH frame count: 5\n Output files prefix: jet\n Transition to IPT (0/1): 1\n"
                    " Velocity equation order (1/2): 2\n Maximum x trajectory coordinate: 0.01\n Primary grid face filename: %s\n"
                    " SPH tracking conversion x coordinate: 0.001\n"   # (IO.cpp:674-679: no tracking unless this lies upstream of max_x)
                    " Boundary mapping filename: tau.bmap\n Restart-data prefix: %s\n" % (mesh, sol))
    out = subprocess.run([BIN, str(para), "--quiet"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=200)
    assert out.returncode == 0 and "Simulation complete!" in out.stdout, out.stdout[-3000:]
    m = re.search(r"Particle tracking: (\d+) particles followed, (\d+) left the mesh or passed the end plane, (\d+) failed", out.stdout)
    assert m, out.stdout[-3000:]
    n, ok, bad = (int(g) for g in m.groups())
    deleted = [int(x) for x in re.findall(r"Deleted particles: (\d+)", (tmp_path / "jet_frame.info").read_text())]
    assert n == deleted[-1] > 20 and ok + bad == n
    text = (tmp_path / "jet_IPT_streaks.dat").read_text()
    assert text.startswith('TITLE = "IPT Streaks"\nVARIABLES = "X", "Y", "Z", "t", "dt", "v", "a", "ptID", "Cell_V", "Cell_Rho", "Cell_ID"\n')
    zones = re.findall(r'ZONE T="Particle (\d+)"\nI= (\d+), J=1, K=1, DATAPACKING=POINT\n((?: .*\n)+)', text)
    assert len(zones) == n and len({z[0] for z in zones}) == n     # one streak per particle
    worth = 0
    for pid, count, body in zones:
        rows = np.array([[float(x) for x in line.split()] for line in body.strip("\n").split("\n")])
        assert rows.shape == (int(count), 11) and (rows[:, 7] == int(pid)).all()
        assert rows[0, 1] > 0.0006 and rows[0, 10] >= 0 and rows[0, 4] == 0.0    # handed over past the plane, inside a cell
        assert (np.diff(rows[:, 3]) >= 0).all()                                 # time runs forward along a streak
        worth += int(count) > 2
    assert worth > 0
