#!/usr/bin/env python
"""Time the UNMODIFIED reference's own Python step loop (BASELINE.md section 3 protocol) and write
profiles/python_reference_timing.json.

Runs ONLY in the build container: it imports /root/reference under the throw-away gym stub of tests/golden/_gymstub
(the GPU box has neither).  bench.py copies the committed figures into `cpu_baseline.python_reference`, with this
provenance, next to the C-port baseline it times live on the box.

  #1 SingleAircraftEnv, randint(0, 9) policy, N = 0 (registered default) and N = 80 (Simulators/config.py), auto-reset
  #2 SingleAircraft2Env, U(-1, 1)^2 actions, N = 80
  #3 SingleAircraftHEREnv step + compute_reward on 4 relabelled goals per transition, N = 80
  #4 Algorithms/MCTS: MCTS(root).best_action(100, 3) on the reset observation of Simulators/SingleAircraftMCTSEnv, N = 80
Each row: one process on one core; rows #1/#2 also multiprocessing.Pool(os.cpu_count()) independent envs, aggregate.
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GCA_REFERENCE", "/root/reference")


def _setup():
    import numpy as np
    np.float = float
    sys.path[:0] = [os.path.join(ROOT, "tests", "golden", "_gymstub"), REF, os.path.join(REF, "Simulators"),
                    os.path.join(REF, "Algorithms", "MCTS")]
    return np


def run_env(args):
    name, n, steps, seed = args
    np = _setup()
    from gym_guidance_collision_avoidance_single.envs import SingleAircraftEnv, SingleAircraft2Env, SingleAircraftHEREnv
    from gym_guidance_collision_avoidance_single.envs.config import Config
    Config.intruder_size = n
    np.random.seed(seed)
    env = {"env": SingleAircraftEnv, "env2": SingleAircraft2Env, "her": SingleAircraftHEREnv}[name]()
    rng = np.random.RandomState(seed + 1)
    env.reset()
    goals = rng.uniform(0, 1, (4, 2))
    t0 = time.perf_counter()
    for _ in range(steps):
        a = int(rng.randint(9)) if name == "env" else rng.uniform(-1, 1, 2)
        ob, r, done, info = env.step(a)
        if name == "her":
            for g in goals:                               # k = 4 relabelled goals (Algorithms/DDPG/DDPG.py:38,308-315)
                env.compute_reward(ob["achieved_goal"], g, None)
        if done:
            env.reset()
    return steps / (time.perf_counter() - t0)


def run_mcts(decisions):
    np = _setup()
    import SingleAircraftMCTSEnv as m
    import config as simcfg
    import nodes_single
    import search_single
    simcfg.Config.intruder_size = 80
    np.random.seed(5)
    env = m.SingleAircraftEnv()
    ob = env.reset()
    t0 = time.perf_counter()
    for _ in range(decisions):
        root = nodes_single.SingleAircraftNode(nodes_single.SingleAircraftState(state=np.asarray(ob, np.float64)))
        search_single.MCTS(root).best_action(100, 3)
    dt = time.perf_counter() - t0
    return {"seconds_per_decision": dt / decisions, "simulations_per_sec": 100 * decisions / dt, "decisions": decisions,
            "unit": "best_action(100 simulations, depth 3), N = 80"}


def main():
    import numpy as np
    cores = os.cpu_count() or 1
    out = {"provenance": "unmodified reference (%s) imported under tests/golden/_gymstub, timed by tools/time_python_reference.py "
                         "in the BUILD container (not on the GPU box: the reference and gym do not travel); random policy, "
                         "auto-reset on done" % REF,
           "python": platform.python_version(), "numpy": np.__version__, "cpu": platform.processor() or platform.machine(),
           "cores_available": cores, "unit": "env-steps/s", "rows": {}}
    try:
        with open("/proc/cpuinfo") as f:
            out["cpu"] = [l.split(":", 1)[1].strip() for l in f if l.startswith("model name")][0]
    except Exception:
        pass
    ctx = mp.get_context("spawn")
    with ctx.Pool(1) as one:
        for name, n, steps in (("env", 0, 20000), ("env", 80, 4000), ("env2", 0, 20000), ("env2", 80, 4000), ("her", 80, 4000)):
            v = one.map(run_env, [(name, n, steps, 0)])[0]
            out["rows"]["%s_n%d" % (name, n)] = {"one_process": v, "steps": steps}
            print(name, n, "1 process: %.1f steps/s" % v, flush=True)
    with ctx.Pool(cores) as pool:
        for name, n, steps in (("env", 80, 1500), ("env2", 80, 1500)):
            t0 = time.perf_counter()
            pool.map(run_env, [(name, n, steps, 10 + c) for c in range(cores)])
            agg = cores * steps / (time.perf_counter() - t0)
            out["rows"]["%s_n%d" % (name, n)]["pool"] = {"value": agg, "processes": cores, "steps_each": steps,
                                                         "note": "wall clock around Pool.map, includes the workers' imports"}
            print(name, n, "Pool(%d): %.1f steps/s" % (cores, agg), flush=True)
    with ctx.Pool(1) as one:
        out["mcts"] = one.map(run_mcts, [3])[0]
        out["mcts"]["provenance"] = out["provenance"]
    print("mcts", out["mcts"])
    out["headline_workload"] = {"row": "env2_n80", "value": out["rows"]["env2_n80"]["one_process"], "cores": 1,
                                "note": "SingleAircraft2Env, N = 80: the reference's own loop for bench.py's workload"}
    with open(os.path.join(ROOT, "profiles", "python_reference_timing.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
