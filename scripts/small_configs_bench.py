import sys, time, tempfile, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import jots_b200 as J
from cases import make_case
d = tempfile.mkdtemp()
for name in ["Reinert_B1", "NL_Reinert_B2", "NL_Reinert_B3", "Danish_Model1"]:
    ini = make_case(name, d)
    drv = J.Driver(ini)
    t0 = time.perf_counter(); n = drv.run(); dt = time.perf_counter() - t0
    st = drv.ctx.stats()
    print("%-16s steps=%d  %.3f s  %.1f steps/s  krylov=%d newton=%d launches=%d  us/launch=%.1f" % (name, n, dt, n / dt, st["krylov_iters"], st["newton_iters"], st["kernel_launches"], 1e6 * dt / st["kernel_launches"]))
