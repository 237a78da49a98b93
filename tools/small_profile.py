"""Where does the time go on the small script configs?  (run plain, and under ncu for a launch list)"""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import scenarios as SC
from fdtd_em_solver_b200 import solver as S
name = sys.argv[1] if len(sys.argv) > 1 else "tutorial"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sc = SC.spec(name)
with contextlib.redirect_stdout(io.StringIO()):
    s = SC.drive(S, sc)
    s._coefficients()
    s.time_vector = np.arange(0, s.end_time, s.dt)
    s._sync_engine(s.time_vector)
e = s._engine
e.run(0, 50)
t0 = time.perf_counter(); e.run(50, n_steps); t1 = time.perf_counter()
print(json.dumps(dict(config=name, N=s.size, steps=n_steps, wall_us_per_step=1e6 * (t1 - t0) / n_steps,
                      device_us_per_step=1e3 * e.stats()["last_run_ms"] / n_steps,
                      tile_rows=e.cfg.tile_rows)))
t0 = time.perf_counter()
for n in range(n_steps):
    e.lib.fdtd_step(e.handle, 50 + n_steps + n)
t1 = time.perf_counter(); e.sync(); t2 = time.perf_counter()
print(json.dumps(dict(enqueue_us_per_step=1e6 * (t1 - t0) / n_steps, total_us_per_step=1e6 * (t2 - t0) / n_steps)))
t0 = time.perf_counter(); snaps = e.run(50 + 2 * n_steps, 500, snapshot_every=5, total_calls=10 ** 6); t1 = time.perf_counter()
print(json.dumps(dict(run500_with_100_snapshots_s=t1 - t0)))
