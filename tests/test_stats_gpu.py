"""Device-side conservation sums vs the CPU restatement of nbody_data::print_statistics
(oracle orc_statistics, itself pinned on the compiled reference in tests/test_oracle.py)."""
import numpy as np
import pytest

from conftest import load_golden_npz
from util import universe

pytestmark = pytest.mark.gpu


def check(stats, want, rel=1e-12):
    for key in ("P", "L", "C"):
        scale = max(1.0, float(np.abs(want[key]).max()))
        assert np.abs(stats[key] - want[key]).max() <= 1e-11 * scale, key
    assert stats["Ekin"] == pytest.approx(want["Ekin"], rel=rel)
    assert stats["Epot"] == pytest.approx(want["Epot"], rel=rel)


@pytest.mark.parametrize("devices", ["0", "0,0"])
@pytest.mark.parametrize("tag", ["g1_n256", "g1_n2048"])
def test_statistics_vs_oracle(oracle64, tag, devices):
    from nbody_b200 import Engine
    g = load_golden_npz(tag)
    with Engine(devices=devices) as e:
        assert e.init(g["y"], g["mass"])
        s = e.statistics()
        check(s, oracle64.statistics(g["y"], g["mass"]))
        lin = e.statistics(with_energy=False)
        assert lin["Epot"] == 0 and lin["Ekin"] == s["Ekin"]
        assert e.statistics(y="not a buffer") is None


def test_statistics_close_pairs_are_skipped(oracle64):
    """potential_energy returns 0 for r2 < MinDistance (nbody_data.cpp:46-55), which also removes self pairs."""
    from nbody_b200 import Engine
    y = np.zeros(6 * 4)
    y[0:4] = [0.0, 0.0, 5e-5, 1.0]
    y[12:16] = [1.0, -1.0, 0.5, 0.25]
    m = np.array([1.0, 2.0, 3.0, 4.0])
    with Engine() as e:
        assert e.init(y, m)
        check(e.statistics(), oracle64.statistics(y, m))


def test_statistics_n16384_energy_and_ranks_free_properties(oracle64):
    n = 16384
    y, m = universe(n)
    from nbody_b200 import Engine
    with Engine() as e:
        assert e.init(y, m)
        s = e.statistics()
    check(s, oracle64.statistics(y, m), rel=1e-11)
    assert s["Epot"] < 0 < s["Ekin"]


def test_statistics_fp32_build(oracle64):
    from nbody_b200 import Engine
    g = load_golden_npz("g1_n2048", "f32")
    with Engine(precision="f32") as e:
        assert e.init(g["y"], g["mass"])
        s = e.statistics()
    want = oracle64.statistics(g["y"].astype(np.float64), g["mass"].astype(np.float64))
    assert s["Epot"] == pytest.approx(want["Epot"], rel=1e-6)
    assert s["Ekin"] == pytest.approx(want["Ekin"], rel=1e-6)
