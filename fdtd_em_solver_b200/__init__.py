"""fdtd_em_solver_b200 -- B200-native 2D Yee FDTD time-stepper behind the Python API of
jmanchuck/FDTD-EM-Solver.  `solver` and `analyser` mirror the reference modules of the
same names; `engine` is the thin ctypes layer over csrc/ (libfdtd2d.so, sm_100a)."""
__all__ = ["solver", "analyser", "engine", "slab", "synthetic"]
