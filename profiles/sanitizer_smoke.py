import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
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
for mode in (0, 4, 32):                # walk variants: two / four / one target(s) per lane
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
