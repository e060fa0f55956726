"""Profile (cProfile) of cm.poincare_map(0.7).compute("p3") with SeedingOptions(n_seeds=N) under hiten_b200.install():
where the host time of the drop-in goes.  usage: prof_cm_dropin.py [N]   (needs a GPU and oracle/_ref)"""
import cProfile, os, pstats, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
import _refenv; _refenv.enable()
from hiten import System
import hiten_b200
from hiten.algorithms.poincare.centermanifold.options import CenterManifoldMapOptions
from hiten.algorithms.poincare.core.options import IterationOptions, SeedingOptions
from hiten.algorithms.types.options import IntegrationOptions, WorkerOptions
system = System.from_bodies("earth", "moon")
cm = system.get_libration_point(1).get_center_manifold(degree=6); cm.compute()
hiten_b200.install(cm_seeds_from_options=True)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for rep in range(2):
    pm = cm.poincare_map(energy=0.7)
    pm.dynamics.clear(); pm.dynamics.reset()
    opts = CenterManifoldMapOptions(integration=IntegrationOptions(dt=0.01, order=4, c_omega_heuristic=20, max_steps=2000),
        iteration=IterationOptions(n_iter=1), seeding=SeedingOptions(n_seeds=N), workers=WorkerOptions(n_workers=1))
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
    pm.compute(section_coord="p3", options=opts)
    pr.disable(); print("rep", rep, "wall", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
