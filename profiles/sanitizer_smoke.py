import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['NBREF_QUIET']='1'
import numpy as np
from nbody_b200 import Engine
rng = np.random.RandomState(0)
n = 2048
y = rng.uniform(-50, 50, 6*n); m = rng.uniform(0.1, 2, n)
for opts in ((), (("direct_symmetric",1),("direct_sym_tile",512)), (("direct_symmetric",1),("direct_sym_tile",256),("direct_sym_shape",0))):
    with Engine(devices="0,0") if not opts else Engine() as e:
        for k,v in opts: e.set_option(k,v)
        assert e.init(y, m)
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0, e.get_y(), f)
        ks = e.create_buffers(e.get_y().size(), 3)
        for k in ks: e.copy_buffer(k, f)
        e.fmaddn(f, e.get_y(), ks, np.array([0.1, 0.0, 0.3]))
        e.fmaddn_corr(f, ks[0], ks[1:], np.array([0.5, 0.25]))
        print(opts, e.fmaxabs(f))
with Engine(devices="0,0") as e:       # symmetric tiles across lanes: peer-sum kernel
    e.set_option("direct_symmetric", 1); e.set_option("direct_sym_tile", 512)
    assert e.init(y, m)
    f = e.create_buffer(e.get_y().size()); g = e.create_buffer(e.get_y().size())
    e.fcompute(0, e.get_y(), f); e.fcompute(0, f, g); e.fcompute(0, e.get_y(), g)
    print("sym lanes", e.last_direct_path(), e.fmaxabs(g))
with Engine(kind="bh", devices="0,0") as e:
    assert e.init(y, m)
    f = e.create_buffer(e.get_y().size())
    e.fcompute(0, e.get_y(), f)
    print("bh", e.fmaxabs(f))
n = 4096
y = rng.uniform(-50, 50, 6*n); m = rng.uniform(0.1, 2, n)
with Engine(kind="bh") as e:
    assert e.init(y, m)
    f = e.create_buffer(e.get_y().size())
    e.fcompute(0, e.get_y(), f)
    print("bh4096", e.fmaxabs(f))
for mode in (0, 2, 4, 32):             # walk variants: grouped (default) / two / four / one target(s) per lane
    with Engine(kind="bh") as e:
        e.set_option("walk_mode", mode)
        assert e.init(y, m)
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0, e.get_y(), f)
        print("bh4096 walk_mode", mode, e.fmaxabs(f))
pos = np.ascontiguousarray(y.reshape(6, n)[0:3].T); vel = np.ascontiguousarray(y.reshape(6, n)[3:6].T)
for dev in ("0", "0,0"):               # body transposes + solver steps replayed as graphs (two segments around fmaxabs)
    with Engine(devices=dev) as e:
        assert e.init_bodies(pos, vel, m)
        e.set_option("step_graph", 1)
        dy = e.create_buffer(e.get_y().size())
        for _ in range(5):
            e.fcompute(0, e.get_y(), dy)
            err = e.fmaxabs(dy)
            e.fmadd_inplace(e.get_y(), dy, 1e-3)
            e.advise_time(1e-3)
        p, v = e.get_bodies()
        print("bodies + step graph", dev, e.step_graph_stats(), err, float(np.abs(p).max()))
n = 2048
y = rng.uniform(-50, 50, 6*n); m = rng.uniform(0.1, 2, n)
for prec in ("f64", "f32"):            # single-launch kernel for small systems; 2-warp / 1-warp symmetric tiles of 256 bodies
    for opts in ((), (("direct_symmetric", 1), ("direct_sym_tile", 256)), (("direct_symmetric", 1), ("direct_sym_tile", 512))):
        with Engine(precision=prec) as e:
            for k, v in opts: e.set_option(k, v)
            assert e.init(y, m)
            f = e.create_buffer(e.get_y().size())
            e.fcompute(0, e.get_y(), f)
            print(prec, opts, "path", e.last_direct_path(), e.fmaxabs(f))
with Engine() as e:                    # step table: two alternating steps, both captured and replayed
    assert e.init(y, m)
    e.set_option("step_graph", 1)
    dy = e.create_buffer(e.get_y().size())
    for i in range(10):
        e.fcompute(0, e.get_y(), dy)
        e.fmadd_inplace(e.get_y(), dy, 1e-3 if i % 2 else 5e-4)
        e.advise_time(1e-3)
    print("alternating steps", e.step_graph_stats(), e.fmaxabs(e.get_y()))

# round 2: grouped walk with counters (STATS instantiation) and in the FP32 build; several shards; a lattice with
# knife-edge cells (the FP64 re-evaluation path); symmetric tiles whose first pass meets close pairs (clamp redo)
n = 4096
y = rng.uniform(-50, 50, 6*n); m = rng.uniform(0.1, 2, n)
for prec, dev in (("f64", "0"), ("f32", "0"), ("f64", "0,0,0,0")):
    with Engine(kind="bh", precision=prec, devices=dev) as e:
        assert e.init(y, m)
        f = e.create_buffer(e.get_y().size())
        e.bh_walk_stats(True)
        e.fcompute(0, e.get_y(), f)
        print("grouped walk", prec, dev, e.bh_walk_stats(False), e.bh_walk_profile()["max_stack"], e.fmaxabs(f))
g = np.arange(16, dtype=np.float64)
pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
yl = np.concatenate([pos[:, 0], pos[:, 1], pos[:, 2], np.zeros(3 * 4096)])
with Engine(kind="bh", distance_to_node_radius_ratio=2.0) as e:
    assert e.init(yl, np.ones(4096))
    f = e.create_buffer(e.get_y().size())
    e.bh_walk_stats(True)
    e.fcompute(0, e.get_y(), f)
    print("lattice", e.bh_walk_profile()["unsure_lane_items"], e.fmaxabs(f))
n = 4096
y = rng.uniform(-50, 50, 6*n); m = rng.uniform(0.1, 2, n)
yy = y.reshape(6, n)
yy[0:3, 3000] = yy[0:3, 10]; yy[0, 3000] += 1e-6      # a close pair across tiles, and a coincident one
yy[0:3, 2000] = yy[0:3, 20]
for shape in (0, 1):
    with Engine() as e:
        e.set_option("direct_symmetric", 1); e.set_option("direct_sym_tile", 512); e.set_option("direct_sym_shape", shape)
        assert e.init(yy.reshape(-1), m)
        f = e.create_buffer(e.get_y().size())
        e.fcompute(0, e.get_y(), f)
        print("sym redo shape", shape, e.last_direct_path(), e.fmaxabs(f))
