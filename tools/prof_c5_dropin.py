"""Profile (cProfile) of the reference's heteroclinic example (examples/heteroclinic_connection.py:29-63) under
hiten_b200.install() with a finer manifold step: where the host time of the drop-in goes when the tubes get large.
usage: prof_c5_dropin.py [step]   (needs a GPU and oracle/_ref)"""
import cProfile, os, pstats, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
import _refenv; _refenv.enable()
from hiten.algorithms.connections import ConnectionPipeline
from hiten.algorithms.connections.config import ConnectionConfig
from hiten.algorithms.connections.options import ConnectionOptions
from hiten.algorithms.poincare import SynodicMapConfig
from hiten.system import System
import hiten_b200
step = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
system = System.from_bodies("earth", "moon"); mu = system.mu
l1, l2 = system.get_libration_point(1), system.get_libration_point(2)
hiten_b200.install()
t0 = time.perf_counter()
halo_l1 = l1.create_orbit('halo', amplitude_z=0.5, zenith='southern'); halo_l1.correct(); halo_l1.propagate()
halo_l2 = l2.create_orbit('halo', amplitude_z=0.3663368, zenith='northern'); halo_l2.correct(); halo_l2.propagate()
print("orbits", time.perf_counter() - t0)
for rep in range(2):
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    manifold_l1 = halo_l1.manifold(stable=True, direction='positive')
    manifold_l1.compute(integration_fraction=0.9, step=step, show_progress=False)
    manifold_l2 = halo_l2.manifold(stable=False, direction='negative')
    manifold_l2.compute(integration_fraction=1.0, step=step, show_progress=False)
    t1 = time.perf_counter()
    conn = ConnectionPipeline.with_default_engine(config=ConnectionConfig(
        section=SynodicMapConfig(section_axis="x", section_offset=1 - mu, plane_coords=("y", "z")), direction=-1))
    result = conn.solve(manifold_l1, manifold_l2, options=ConnectionOptions(delta_v_tol=1, ballistic_tol=1e-8, eps2d=1e-3))
    pr.disable()
    print("rep", rep, "tubes", t1 - t0, "connections", time.perf_counter() - t1, "n", len(result) if hasattr(result, "__len__") else result)
pstats.Stats(pr).sort_stats("cumulative").print_stats(50)
