"""Reference-order mode on small scenes: step time against the number of blocks of the solve kernels (NB2_REF_BLOCKS).
Run under gpurun."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from nphysics_b200 import abi, scenes
    from nphysics_b200.solver import Solver
    sc = {"pyramid3": scenes.pyramid3, "wall3": scenes.wall3, "boxes 6^3": lambda: scenes.boxes3(6, 6, 6),
          "boxes 20x10x20": lambda: scenes.boxes3(20, 10, 20)}[sys.argv[2]]()
    mode = abi.MODE_REFERENCE_ORDER if sys.argv[3] == "ref" else abi.MODE_COLOURED
    s = Solver(0); s.set_params(sc.params); s.upload_bodies(sc.bodies)
    s.upload_colliders(scenes.scene_colliders(sc)); s.detect_pairs(scenes.LINEAR_PREDICTION)
    for _ in range(10):
        s.generate_manifolds(); s.step(mode)
    s.enable_timers(True)
    acc = {}
    for _ in range(20):
        s.generate_manifolds(); s.step(mode)
        for k, v in s.get_timers().items(): acc[k] = acc.get(k, 0.0) + v / 20
    st = s.get_stats()
    print(json.dumps({"scene": sys.argv[2], "mode": sys.argv[3], "blocks": os.environ.get("NB2_REF_BLOCKS", "auto"),
                      "step_ms": round(acc["step"], 4), "velocity_ms": round(acc["velocity_resolution"], 4),
                      "position_ms": round(acc["position_resolution"], 4), "assembly_ms": round(acc["assembly"], 4),
                      "phases_v": int(st["n_phases_velocity"]), "phases_p": int(st["n_phases_position"])}))
else:
    for scene in ("pyramid3", "wall3", "boxes 6^3", "boxes 20x10x20"):
        for blocks in ("0", "1", "2", "4", "8"):
            env = dict(os.environ, NB2_REF_BLOCKS=blocks)
            subprocess.run([sys.executable, __file__, "child", scene, "ref"], env=env)
        subprocess.run([sys.executable, __file__, "child", scene, "col"])
