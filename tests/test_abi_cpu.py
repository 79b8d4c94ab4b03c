"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and its host-side settings path (defaults, para parser, Set_Values) agrees with the oracle's restatement."""
import os

import pytest

from fjsph_b200 import _lib, engine as eng
from oracle import oracle as orc


@pytest.fixture(scope="module", autouse=True)
def _built():
    _lib.build()


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in L.fjsph_version()


@pytest.mark.parametrize("dim", [2, 3])
def test_set_values_matches_oracle(dim):
    kw = dict(particle_step=0.0015, speed_sound=100.0, rho_rest=810.0, mu=0.000142, sig=0.0256, press_pipe=50.0,
              delta_t_min=1e-9, i_interp_fac=0.25, pressure_rel=dim - 2)
    a = eng.params_to_dict(eng.default_params(dim, **kw))
    b = orc.params_to_dict(orc.default_params(dim, **kw))
    assert a.keys() == b.keys()
    for k in a:
        assert a[k] == b[k], k


def test_set_values_rejects_missing_spacing():
    p = _lib.FjsphParams()
    L = _lib.lib()
    assert L.fjsph_default_params(p, 3) == 0
    assert L.fjsph_set_values(p) != 0
    assert b"spacing" in L.fjsph_last_error()
    assert L.fjsph_default_params(p, 4) != 0


def test_read_para_droplet_deck(tmp_path):
    deck = tmp_path / "para3D"
    deck.write_text(
        "\n".join([
            "      Input fluid definition filename: fluid_3D.bmap",
            "   Input boundary definition filename: boundary.bmap",
            "                   Reference velocity: 21.55",
            "                   Reference pressure: 100000",
            "                Reference temperature: 298",
            "                    Reference density: 1.1025",
            "#                    Reynolds number: 3.231E+06",
            "       Sutherland reference viscosity: 1.716e-05",
            "          Reference dispersed density: 810",
            "        Reference dispersed viscosity: 0.000142",
            "            Reference surface tension: 0.0256",
            "# SPH parameters --------------------: -",
            "              SPH frame time interval: 1e-3",
            "                 SPH maximum timestep: 1",
            "                 SPH minimum timestep: 1e-9",
            "                    SPH CFL condition: 0.9   # trailing comment",
            "SPH stable CFL count iteration factor: 0.76",
            "                  SPH initial spacing: 0.0015",
            "                 SPH aerodynamic case: Gissler",
            "                SPH starting pressure: 000",
            "                   SPH speed of sound: 100",
            "      SPH artificial viscosity factor: 0.05",
            "              SPH freestream velocity: 0,21.55,0",
            "               SPH integration solver: Runge-Kutta",
            "                   SPH gravity vector: 0, 0, -1.5",
            "",
        ])
    )
    p, fluid, bound = eng.read_para(str(deck))
    assert (fluid, bound) == ("fluid_3D.bmap", "boundary.bmap")
    assert p.particle_step == 0.0015 and p.cfl == 0.9 and p.subits_factor == 0.76
    assert p.acase == 1 and p.solver_type == 1
    assert list(p.v_inf) == [0.0, 21.55, 0.0] and list(p.grav) == [0.0, 0.0, -1.5]
    assert p.rho_rest == 810 and p.mu == 0.000142 and p.mu_g == 1.716e-05 and p.sig == 0.0256
    assert p.p_ref == 100000 and p.rho_g == 1.1025 and p.delta_t_min == 1e-9 and p.delta_t == 1e-9
    assert p.frame_time_interval == 1e-3 and p.visc_alpha == 0.05
    assert p.H == 0.003 and p.B == pytest.approx(810 * 100.0**2 / 7)
    bad = tmp_path / "bad"
    bad.write_text("SPH initial spacing: 0.1\nSPH integration solver: Leapfrog\n")
    with pytest.raises(_lib.FjsphError):
        eng.read_para(str(bad))
    with pytest.raises(_lib.FjsphError):
        eng.read_para(str(tmp_path / "missing"))


def test_engine_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = eng.default_params(3, particle_step=1e-3)
    with pytest.raises(_lib.FjsphError) as ei:
        eng.Engine(p, 10)
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    root = os.path.join(_lib.ROOT, "fjsph_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.replace("the CPU oracle", "").replace("CPU oracle", ""), os.path.join(dirpath, f)
