#!/usr/bin/env python3
"""Fixtures the reference's OWN tests hold for the hot path, restated as arrays -- build container only.

  ref_reproducibility.npz   tests/test_data/reproducibility_data.py:5-140 (COMPLICATED_ENVIRONMENT, the golden
      trajectory of tests/test_integration.py:66-93): constructor arguments, seed, actions, observations,
      rewards, terminated, truncated, exactly as the reference file holds them. The fork's driver profiles and
      rule engine made the file partly stale: the `traffic` plane differs from reset on and
      from step 18 the rule engine brakes; `live_steps` = 18 and `live_planes` name what the CURRENT reference
      code still reproduces (checked here against the live reference with ignore_traffic_collisions=True).
  ref_next_subgoal_direction.npz   the table of tests/test_environment.py:606-640 on the reference's
      tests/test_data/map_with_all_directions.json (map plan included), with the answers the CURRENT reference
      code gives (13 of the 15 rows equal the file's; the two `-1` rows return a compass index since the fork's
      fallback at environment.py:1470-1502) next to the file's.

    python tests/golden/make_reference_fixtures.py
"""
import ast
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PGTG_REFERENCE", "/root/reference")


def nsd_table_from_reference_test():
    """The parametrize list of TestObservation.test_next_subgoal_direction, read from the reference's test file."""
    src = open(os.path.join(REF, "tests", "test_environment.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "test_next_subgoal_direction":
            for dec in node.decorator_list:
                if isinstance(dec, ast.Call) and getattr(dec.func, "attr", "") == "parametrize":
                    return ast.literal_eval(dec.args[1])
    raise RuntimeError("table not found")


def main():
    from oracle import ref_runner

    environment = ref_runner.import_reference()
    sys.path.insert(0, REF)
    from tests.test_data import reproducibility_data as rd

    g = rd.COMPLICATED_ENVIRONMENT
    kw = dict(g["environment_arguments"])
    keys = list(kw["features_to_include_in_observation"])
    obs = g["observation_list"]
    T = len(g["action_list"])
    maps = np.stack([np.stack([np.asarray(o["map"][k], np.int8) for k in keys]) for o in obs])  # [T+1, C, 9, 9]
    pos = np.stack([np.asarray(o["position"], np.int32) for o in obs])
    vel = np.stack([np.asarray(o["velocity"], np.int32) for o in obs])
    # which part is live: replay the current reference (collisions off: the fork's traffic crashes the agent at step 8)
    live_planes = [k for k in keys if k != "traffic"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = environment.PGTGEnv(**{**kw, "ignore_traffic_collisions": True})
        o, _ = env.reset(seed=g["seed"])
        live = 0

        def same(o, t):
            return (all(np.array_equal(np.asarray(o["map"][k]), maps[t, keys.index(k)]) for k in live_planes)
                    and np.array_equal(o["position"], pos[t]) and np.array_equal(o["velocity"], vel[t]))

        assert same(o, 0)
        for t in range(T):
            o, r, term, trunc, _ = env.step(g["action_list"][t])
            if not (same(o, t + 1) and r == g["reward_list"][t] and term == g["terminated_list"][t]):
                break
            live = t + 1
    print("golden trajectory: live for reset +", live, "steps; planes", live_planes)
    meta = dict(kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, seed=g["seed"], keys=keys, live_steps=live, live_planes=live_planes,
                source="tests/test_data/reproducibility_data.py:5-140, tests/test_integration.py:66-93")
    np.savez_compressed(os.path.join(HERE, "ref_reproducibility.npz"), obs_map=maps, obs_position=pos, obs_velocity=vel,
                        actions=np.asarray(g["action_list"], np.int32), reward=np.asarray(g["reward_list"], np.float64),
                        terminated=np.asarray(g["terminated_list"], np.uint8), truncated=np.asarray(g["truncated_list"], np.uint8),
                        meta=np.frombuffer(json.dumps(meta).encode(), np.uint8))

    table = nsd_table_from_reference_test()
    plan = json.load(open(os.path.join(REF, "tests", "test_data", "map_with_all_directions.json")))
    live_answers = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for (tx, ty), _want in table:
            env = environment.PGTGEnv(map_path=os.path.join(REF, "tests", "test_data", "map_with_all_directions"), use_next_subgoal_direction=True)
            env.reset(seed=0)
            env.position = np.array([tx * 9 + 4, ty * 9 + 4])
            o, _, _, _, _ = env.step(4)
            live_answers.append(int(o["next_subgoal_direction"]))
    file_answers = [w for _, w in table]
    print("next_subgoal_direction: file", file_answers, "\n                 current code", live_answers)
    meta = dict(map_name="map_with_all_directions.json", plan=plan, source="tests/test_environment.py:606-640, tests/test_data/map_with_all_directions.json")
    np.savez_compressed(os.path.join(HERE, "ref_next_subgoal_direction.npz"), tile=np.asarray([t for t, _ in table], np.int32),
                        file_answer=np.asarray(file_answers, np.int32), live_answer=np.asarray(live_answers, np.int32),
                        meta=np.frombuffer(json.dumps(meta).encode(), np.uint8))


if __name__ == "__main__":
    main()
