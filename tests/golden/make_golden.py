"""Regenerates tests/golden/ from the reference (run in the build container only,
where /root/reference is mounted and oracle/_ref has been built):

    python tests/golden/make_golden.py

* the reference's own solver golden vectors (test/data/initial_state.txt and the
  19 end states, tolerance 1e-12 in test/solver/test_nbody_solver.cpp:81-84) and
  loader fixtures (zeno_ascii/zeno_table) are copied verbatim -- they are test
  DATA, not sources;
* make_universe states (libstdc++ mt19937_64 stream, nbody_data.cpp:263-324) and
  the reference engines' outputs on them are captured as .npz so the GPU box,
  which has no /root/reference, can still compare against the real reference.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("NBREF_QUIET", "1")

from oracle import refharness as R  # noqa: E402

REF_DATA = "/root/reference/test/data"


def copy_reference_vectors():
    for name in sorted(os.listdir(REF_DATA)):
        if name.endswith(".txt"):
            shutil.copyfile(os.path.join(REF_DATA, name), os.path.join(HERE, name))


def universe_case(lib, stars, tag, engines, bh_ratios=()):
    d = R.Data(lib).make_universe(stars)
    y, m = d.export()
    out = {"y": y, "mass": m}
    for name, kw in engines.items():
        e = R.Engine(lib, **kw)
        assert e.init(d)
        out["f_" + name] = e.fcompute_y()
        e.close()
    for ratio in bh_ratios:
        h = R.Heap(lib)
        h.build(y, m, ratio)
        xyzr, hm, body = h.export()
        key = ("%g" % ratio).replace(".", "p")
        out["tree_xyzr_" + key] = xyzr
        out["tree_mass_" + key] = hm
        out["tree_body_" + key] = body.astype(np.int32)
        h.close()
    d.close()
    np.savez_compressed(os.path.join(HERE, "%s_%s.npz" % (tag, lib.precision)), **out)
    print(tag, lib.precision, "N =", m.size)


def main():
    copy_reference_vectors()
    for prec in ("f64", "f32"):
        lib = R.load(prec)
        direct = {"simple": dict(engine="simple"), "openmp": dict(engine="openmp"), "block": dict(engine="block")}
        bh = lambda r: dict(engine="simple_bh", distance_to_node_radius_ratio=r, traverse_type="nested_tree",
                            tree_layout="heap_stackless")
        # N = 128 and 256: the fixture sizes of test_nbody_engine.cpp:660-669
        for stars, tag in ((64, "g1_n128"), (128, "g1_n256")):
            eng = dict(direct)
            eng["bh_1e8"] = bh(1e8)
            eng["bh_3p1623"] = bh(3.1623)
            eng["bh_10"] = bh(10)
            universe_case(lib, stars, tag, eng, bh_ratios=(3.1623, 10))
        # C1 of BASELINE.json: --stars_count=1024 -> N = 2048
        eng = dict(direct)
        eng["bh_10"] = bh(10)
        eng["bh_1"] = bh(1)
        universe_case(lib, 1024, "g1_n2048", eng, bh_ratios=(10,))


if __name__ == "__main__":
    main()
