"""The reference's simulation flow with both overlay patches applied (oracle/_ref/nbody_sim_f64: its own sources,
factories and solver loop; integration/nbody_engines.patch + nbody_data.patch), run as configured in BASELINE C1:
--solver=rk4 --stars_count=1024 --initial_type=G1, max_time=1, check_step=0.1, check_list=PLVE. The --check_step report
of --engine=b200 (conservation sums on the device, nbody_engine_b200::statistics) is compared LINE BY LINE with the
report of --engine=openmp (host sums) over the whole run."""
import os

import pytest

from sim_util import run_sim, sim_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(sim_path("f64")):
        pytest.skip("oracle/_ref/nbody_sim_f64 not built (needs /root/reference)")
    return sim_path("f64")


C1 = dict(solver="rk4", stars_count=1024, initial_type="G1", max_time=1, check_step=0.1, check_list="PLVE")


def assert_reports_agree(a, b, lines=10):
    assert len(a) == len(b) == lines
    for ra, rb in zip(a, b):
        # same steps, same times, same number of fcompute calls at every report
        assert (ra["step"], ra["t"], ra["CC"], ra["SC"]) == (rb["step"], rb["t"], rb["CC"], rb["SC"])
        # The report prints drifts in units of the initial value. For C1 over t = 1 they sit at the rounding floor of the
        # sums themselves (dE ~ 2e-10 is 1e-12 of |E| after 2048^2 potential terms), so host (Kahan, body order) and
        # device (blocked) sums agree to that floor, not to the report's four digits:
        assert rb["dE"] == pytest.approx(ra["dE"], rel=2e-3, abs=2e-12), (ra, rb)
        assert rb["dL"] == pytest.approx(ra["dL"], rel=2e-2, abs=1e-13), (ra, rb)
        assert rb["Vcm"] == pytest.approx(ra["Vcm"], rel=1e-6, abs=1e-12), (ra, rb)
        # dP is relative to a total momentum that all but cancels in G1: the report amplifies the last bits of the sum
        # about a million times, so only its size can agree (Barnes-Hut does not conserve momentum: there dP is large and
        # must agree to the report's digits)
        assert rb["dP"] == pytest.approx(ra["dP"], rel=2e-3, abs=2e-9), (ra, rb)


def test_c1_check_step_report_matches_openmp_line_by_line(sim):
    cpu, cpu_sum, _ = run_sim(engine="openmp", **C1)
    gpu, gpu_sum, _ = run_sim(engine="b200", **C1)
    assert cpu_sum["steps"] == gpu_sum["steps"] == 100 and cpu_sum["fcompute_calls"] == gpu_sum["fcompute_calls"] == 400
    assert gpu_sum["engine"] == "nbody_engine_b200"
    assert_reports_agree(cpu, gpu)


def test_c1_report_with_step_graphs_off_and_two_lanes(sim):
    base, _, _ = run_sim(engine="b200", **C1)
    for extra in (dict(step_graph=0), dict(device="0,0")):
        other, summary, _ = run_sim(engine="b200", **dict(C1, **extra))
        assert summary["fcompute_calls"] == 400
        assert_reports_agree(base, other)


def test_c4_style_barnes_hut_run_reports(sim):
    """b200_bh through the patched factory and the patched report (smaller than C4 so that the CPU side finishes):
    euler, 8192 bodies, ratio 10, against the reference's simple_bh engine with the same tree layout."""
    cfg = dict(solver="euler", stars_count=4096, max_time=0.1, check_step=0.01, check_list="PLVE", max_step=0.01,
               distance_to_node_radius_ratio=10, tree_layout="heap_stackless")
    cpu, _, _ = run_sim(engine="simple_bh", traverse_type="nested_tree", **cfg)
    gpu, gs, _ = run_sim(engine="b200_bh", **cfg)
    assert gs["engine"] == "nbody_engine_b200_bh"
    assert_reports_agree(cpu, gpu, lines=11)      # 10 x 0.01 < 0.1 in floating point: the reference's loop takes an 11th step
