"""Cost of ActivationManager::update on the 100k-box pile and of a step once the pile sleeps (run under gpurun)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nphysics_b200 import abi, scenes  # noqa: E402
from nphysics_b200.solver import Solver  # noqa: E402

sc = scenes.boxes3(50, 40, 50)
m, c = scenes.ContactGenerator(sc).generate()
p = abi.default_params()
p["max_velocity_iterations"] = 10
p["max_position_iterations"] = 5
s = Solver(0)
s.set_params(p)
s.upload_bodies(sc.bodies)
s.upload_manifolds(m, c)
n = len(sc.bodies)


def timed(fn, reps):
    s.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    s.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


act = abi.new_activation(n)                      # everybody awake (4 x threshold)
s.upload_activation(act)
for _ in range(20):
    s.step(abi.MODE_COLOURED)
t_update_awake = timed(lambda: s.update_activation(0.01), 20)
t_step_awake = timed(lambda: s.step(abi.MODE_COLOURED), 20)
act["energy"][:] = 0.5 * abi.DEFAULT_SLEEP_THRESHOLD  # everybody below the threshold: the island goes to sleep
rest = np.zeros(n, dtype=abi.body_state_dtype)       # ... once it is at rest (a 40-high pile still creeps at 10 iterations)
rest["position"] = sc.bodies["position"]
s.upload_body_states(rest)
s.upload_activation(act)
s.update_activation(0.01)
a = s.download_activation()
asleep = int((a["energy"][1:] == 0).sum())
for _ in range(3):
    s.step(abi.MODE_COLOURED)
t_step_asleep = timed(lambda: s.step(abi.MODE_COLOURED), 20)
t_update_asleep = timed(lambda: s.update_activation(0.01), 20)
st = s.get_stats()
print({"bodies": n - 1, "asleep_after_update": asleep, "update_ms_awake": round(t_update_awake, 4),
       "step_ms_awake": round(t_step_awake, 4), "step_ms_asleep": round(t_step_asleep, 4),
       "update_ms_asleep": round(t_update_asleep, 4),
       "rows_asleep": int(st["n_rows_two_body"]) + int(st["n_rows_ground"])})
