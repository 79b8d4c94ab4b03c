"""The command-line driver (csrc/fjsph_run.cpp = FJSPH's main(), FJSPH.cpp:29-356, on the C ABI): a deck runs through
the frame loop, writes the frame table, the ASCII frames and the restart file, and a run resumed from the restart file
ends where the uninterrupted one does."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fjsph_b200", "bin", "fjsph_b200_run")


def frame(path):
    return np.loadtxt(path, skiprows=3)


def test_driver_runs_a_deck_and_resumes(tmp_path):
    if not os.path.exists(BIN):
        pytest.fail("driver not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    for f in ("droplet3d.para", "droplet3d_fluid.bmap", "droplet3d_boundary.bmap"):
        shutil.copy(os.path.join(ROOT, "tests", "decks", f), tmp_path / f)
    para = tmp_path / "droplet3d.para"
    para.write_text(para.read_text().replace("SPH frame time interval: 0.001", "SPH frame time interval: 4e-5")
                    + "\n SPH frame count: 4\n Output files prefix: full\n")
    run = lambda *a: subprocess.run([BIN, str(para), "--quiet", *a], cwd=tmp_path, stdout=subprocess.PIPE,
                                    stderr=subprocess.STDOUT, text=True, timeout=120)
    out = run()
    assert out.returncode == 0 and "Simulation complete!" in out.stdout, out.stdout[-2000:]
    info = (tmp_path / "full_frame.info").read_text()
    assert info.count("Frame:") == 4 and "Total Points: 4166 Boundary Points: 0 Fluid Points: 4166" in info
    frames = sorted(p.name for p in tmp_path.glob("full_frame_*.dat"))
    assert frames == ["full_frame_%05d.dat" % k for k in range(4)]
    last = frame(tmp_path / "full_frame_00003.dat")
    assert last.shape == (4166, 12) and np.isfinite(last).all()
    # the droplet sits in a 21.55 m/s cross flow along +y: the Gissler drag has started to move it
    assert last[:, 4].mean() > 0.0
    # resume from the restart file written after frame 2 of a shorter run
    out2 = run("--frames", "3", "--out", "part")
    assert out2.returncode == 0, out2.stdout[-2000:]
    out3 = run("--frames", "4", "--out", "part", "--restart", "part_particles.fjr")
    assert out3.returncode == 0 and "Frame: 3" in out3.stdout, out3.stdout[-2000:]
    resumed = frame(tmp_path / "part_frame_00003.dat")
    assert np.array_equal(resumed[:, 11], last[:, 11])           # same particles, same order
    assert np.abs(resumed[:, :3] - last[:, :3]).max() <= 1e-9 * 0.05  # positions, relative to the droplet radius
    assert (tmp_path / "part_frame.info").read_text().count("Frame:") == 4
    # the same deck coupled to an OpenFOAM case whose solution is the uniform free stream: the containment lookup
    # (FindCell on the device) must hand every particle the free-stream values, i.e. the frames of the constant-velocity run
    from tests.foam_case import write_case

    write_case(tmp_path / "foam", (-0.1013, -0.1007, -0.1011), (0.1009, 0.1003, 0.1017), (6, 7, 5),
               lambda c: (0.0, 21.55, 0.0), lambda c: 100000.0)
    para.write_text(para.read_text() + " OpenFOAM input directory: foam\n OpenFOAM solution directory: 100\n")
    out4 = run("--out", "mesh")
    assert out4.returncode == 0 and "OpenFOAM mesh: 210 cells, 1474 triangles" in out4.stdout, out4.stdout[-2000:]
    coupled = frame(tmp_path / "mesh_frame_00003.dat")
    assert np.array_equal(coupled[:, 11], last[:, 11]) and np.abs(coupled[:, :6] - last[:, :6]).max() == 0.0
    # errors are messages and exit codes, never a crash
    bad = subprocess.run([BIN, str(tmp_path / "nope.para")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert bad.returncode == 1 and "could not open" in bad.stdout


def test_driver_two_ranks_native_nccl(tmp_path):
    """fjsph_b200_run --ranks 2: one process per GPU forked by the driver, the case cut into two x-slabs, ghosts and
    migration by ncclSend / ncclRecv, the step's scalars by ncclAllReduce on device buffers (csrc/comm_nccl.cu) -- no Python
    anywhere.  The particles of the two slabs, put together by id, are the single-GPU run's.  Needs two visible GPUs."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for f in ("droplet3d.para", "droplet3d_fluid.bmap", "droplet3d_boundary.bmap"):
        shutil.copy(os.path.join(ROOT, "tests", "decks", f), tmp_path / f)
    para = tmp_path / "droplet3d.para"
    para.write_text(para.read_text().replace("SPH frame time interval: 0.001", "SPH frame time interval: 4e-5")
                    + "\n SPH frame count: 3\n")
    run = lambda *a: subprocess.run([BIN, str(para), "--quiet", *a], cwd=tmp_path, stdout=subprocess.PIPE,
                                    stderr=subprocess.STDOUT, text=True, timeout=300)
    one = run("--out", "one")
    assert one.returncode == 0, one.stdout[-2000:]
    two = run("--out", "two", "--ranks", "2")
    assert two.returncode == 0 and "NCCL transport:" in two.stdout, two.stdout[-3000:]
    assert " 0 device all-reduces" not in two.stdout  # the residual was reduced on the device
    ref = frame(tmp_path / "one_frame_00002.dat")
    parts = np.concatenate([frame(tmp_path / ("two_r%d_frame_00002.dat" % r)) for r in (0, 1)])
    assert parts.shape == ref.shape
    parts = parts[np.argsort(parts[:, 11])]
    ref = ref[np.argsort(ref[:, 11])]
    assert np.array_equal(parts[:, 11], ref[:, 11])
    # frames are printed with 8 significant digits; the slabs differ from one engine by summation order only
    assert np.abs(parts[:, :3] - ref[:, :3]).max() <= 2e-7 * 0.05
    assert np.abs(parts[:, 3:6] - ref[:, 3:6]).max() <= 1e-5 * max(np.abs(ref[:, 3:6]).max(), 1e-12)
    info = (tmp_path / "two_r0_frame.info").read_text()
    assert info.count("Frame:") == 3 and "Total Points: 4166 Boundary Points: 0 Fluid Points: 4166" in info


def test_driver_runs_the_2d_dam_break(tmp_path):
    """Examples/Dam_2D through the driver: `--dim 2` is the reference's 2D build target (makefile: SIMDIM=2)."""
    if not os.path.exists(BIN):
        pytest.fail("driver not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    for f in ("dam2d.para", "dam2d_fluid.bmap", "dam2d_boundary.bmap"):
        shutil.copy(os.path.join(ROOT, "tests", "decks", f), tmp_path / f)
    para = tmp_path / "dam2d.para"
    para.write_text(para.read_text().replace("SPH frame time interval: 0.01", "SPH frame time interval: 0.002")
                    + "\n SPH frame count: 3\n Output files prefix: dam\n")
    out = subprocess.run([BIN, str(para), "--quiet", "--dim", "2"], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0 and "Simulation complete!" in out.stdout, out.stdout[-2000:]
    info = (tmp_path / "dam_frame.info").read_text()
    assert info.count("Frame:") == 3 and "Total Points: 1288 Boundary Points: 488 Fluid Points: 800" in info
    last = frame(tmp_path / "dam_frame_00002.dat")
    assert last.shape == (1288, 12) and np.isfinite(last).all()
    assert np.abs(last[:, 2]).max() == 0.0 and np.abs(last[:, 5]).max() == 0.0  # z and v_z stay exact zeros
    fluid = last[488:]
    assert fluid[:, 4].mean() < 0.0  # the column has started to fall (gravity along -y)
    # a 3D read of the 2D deck is an error message, not a crash
    bad = subprocess.run([BIN, str(para), "--quiet"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=60)
    assert bad.returncode != 0 and "ERROR" in bad.stdout
