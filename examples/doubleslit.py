"""BASELINE config 2 (reference doubleslit.py): a TF/SF plane pulse travelling right through a reflecting wall with
two slits, N = 319, 31 938 calls (5.2 GB of snapshots at full length: try --time-scale 0.2 first).

    python examples/doubleslit.py --time-scale 0.2
"""
from _common import options, run

opt = options(__doc__, "doubleslit")
from fdtd_em_solver_b200.solver import Solver, C

size = 4
solver = Solver(points_per_wavelength=10, stability=0.1, eps_r_max=1, mu_r_max=1, simulation_size=size,
                simulation_time=10 * size / C * opt.time_scale, dtype=opt.dtype)
solver.add_oscillating_pulse(3e9, (2, 0.1), 6e9, direction="right")
solver.create_material()
for upper_left, lower_right in (((0, 1), (1.62, 1.05)), ((1.76, 1), (2.24, 1.05)), ((2.38, 1), (size, 1.05))):
    solver.set_reflect_square(upper_left, lower_right)          # the wall, leaving two slits open
solver.set_reflect_boundaries(up=False, down=False, left=False, right=False)

loader = run(solver, opt)
if loader is not None:
    behind = loader.create_data_collector((2.0, 3.0))           # on the axis, behind the slits
    print("\n%d snapshots; |h| behind the slits peaks at %.4g" % (len(loader.get_matrix()), abs(behind.data).max()))
