"""BASELINE config 3 (reference lens.py): a sinusoidal TF/SF plane wave focused by a convex eps_r = 5 lens, N = 178,
3559 calls.

    python examples/lens.py
"""
import numpy as np

from _common import options, run

opt = options(__doc__, "lens")
from fdtd_em_solver_b200.solver import Solver

solver = Solver(points_per_wavelength=10, stability=0.1, eps_r_max=5, mu_r_max=1, simulation_size=3,
                simulation_time=2e-8 * opt.time_scale, dtype=opt.dtype)
solver.add_sinusoidal_pulse(location=(1.5, 0.2), omega_0=5e9, direction="right")
material = solver.create_material()
material.set_material_convex((1.5, 1), radius=2.2, thickness=1.2, epsilon_rel=5)
solver.set_reflect_boundaries(up=False, down=False, right=False, left=False)

loader = run(solver, opt)
if loader is not None:
    last = loader.get_matrix()[-1]
    i, j = np.unravel_index(np.argmax(np.abs(last)), last.shape)
    print("\n%d snapshots; strongest field of the last one at (%.2f, %.2f) m" % (len(loader.get_matrix()),
                                                                              i * loader.ds, j * loader.ds))
