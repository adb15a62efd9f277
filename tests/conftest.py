import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def rollout_from_golden(g, prefix="in_"):
    """Rebuild the SyntheticRollout a golden fixture was generated from."""
    from ppo_and_friends_b200.synthetic import SyntheticRollout, make_rollout
    agents = [str(a) for a in g[prefix + "agents"]]
    if (prefix + "make_rollout_json") in g:
        # large fixtures: the synthetic observations / rewards are regenerated from the recorded make_rollout arguments
        # (deterministic numpy generator); what the reference's networks produced on them is read from the fixture
        import json
        kw = json.loads(str(g[prefix + "make_rollout_json"]))
        if "agents" in kw:
            kw["agents"] = tuple(kw["agents"])
        ro = make_rollout(**kw)
        assert np.array_equal(ro.terminated, g[prefix + "terminated"]) and np.array_equal(ro.truncated, g[prefix + "truncated"])
        for a in agents:
            for name in ("raw_actions", "actions", "values", "log_probs", "next_values"):
                getattr(ro, name)[a] = g[f"{prefix}{name}/{a}"]
        return ro
    ro = SyntheticRollout(T=int(g[prefix + "T"]), E=int(g[prefix + "E"]), agents=agents,
                          obs_dim=int(g[prefix + "obs_dim"]), critic_obs_dim=int(g[prefix + "critic_obs_dim"]),
                          act_dim=int(g[prefix + "act_dim"]), n_discrete=int(g[prefix + "n_discrete"]),
                          max_ts_per_ep=int(g[prefix + "max_ts_per_ep"]))
    ro.terminated = g[prefix + "terminated"]
    ro.truncated = g[prefix + "truncated"]
    for a in agents:
        for name in ("obs", "next_obs", "critic_obs", "raw_actions", "actions", "values",
                     "log_probs", "rewards", "next_values"):
            getattr(ro, name)[a] = g[f"{prefix}{name}/{a}"]
    return ro
