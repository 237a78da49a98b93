"""BASELINE config 4 (reference waveguide.py): a point pulse feeding a bent eps_r = 2.7 dielectric waveguide between
two reflecting plates, with an energy probe in the guide, N = 139, 5577 calls.  --mixed-walls reflects at the upper and
left walls and absorbs at the other two.

    python examples/waveguide.py [--mixed-walls]
"""
import sys

from _common import options, run

mixed = "--mixed-walls" in sys.argv
if mixed:
    sys.argv.remove("--mixed-walls")
opt = options(__doc__, "waveguide")
from fdtd_em_solver_b200.solver import Solver, C

size = 4
solver = Solver(points_per_wavelength=10, stability=0.1, eps_r_max=3.5, mu_r_max=1, simulation_size=size,
                simulation_time=4 * size / C * opt.time_scale, dtype=opt.dtype)
solver.set_reflect_boundaries(False, False, False, False)
solver.add_oscillating_pulse(5e8, (0.7, 0), 2e9)
material = solver.create_material()
material.set_fixed_length_waveguide((0.5, 0), 6, 0.1, 0.4, 2.7)
solver.set_reflect_square((0, 0.25), (0.5, 0.35))
solver.set_reflect_square((0.9, 0.25), (4, 0.35))
solver.record_energy((0.7, 0.1))
if mixed:
    solver.set_reflect_boundaries(up=True, down=False, left=True, right=False)

loader = run(solver, opt)
energies = solver.recorded_energies[0]["energies"]
if energies:
    print("\nenergy probe at (0.7, 0.1) m: %d samples, peak %.4g J/m^3" % (len(energies), max(energies)))
