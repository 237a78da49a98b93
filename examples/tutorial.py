"""BASELINE config 1 (reference tutorial.py, repaired as SURVEY.md App. C): one oscillating point pulse in vacuum,
absorbing walls, N = 255, 3184 calls, a snapshot every 5 calls.

    python examples/tutorial.py                 # writes tutorial.npy / tutorial.txt, prints the probe spectrum peak
    python examples/tutorial.py --show          # the reference's realtime view
"""
import numpy as np

from _common import options, run

opt = options(__doc__, "tutorial")
from fdtd_em_solver_b200.solver import Solver

sigma_w, omega_0 = 1e9, 5e9                     # bandwidth and centre frequency of the pulse (rad/s)
solver = Solver(points_per_wavelength=10, stability=0.2, eps_r_max=4, mu_r_max=1, simulation_size=3,
                simulation_time=2.5e-8 * opt.time_scale, dtype=opt.dtype)
solver.add_oscillating_pulse(sigma_w, (0.8, 0.4), omega_0)
solver.create_material()
solver.set_reflect_boundaries(up=False, down=False, left=False, right=False)

loader = run(solver, opt)
if loader is not None:
    probe = loader.create_data_collector((1.5, 1.5))
    probe.fft()
    peak = probe.fft_frequencies[int(np.argmax(probe.fft_amplitude))]
    print("\n%d snapshots of %dx%d; spectrum at (1.5, 1.5) m peaks at %.3g rad/s (pulse centre %.3g)"
          % (len(loader.get_matrix()), loader.size, loader.size, peak, omega_0))
