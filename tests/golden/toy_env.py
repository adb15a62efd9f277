"""A tiny deterministic gymnasium-style environment (the interface the reference's SingleAgentGymWrapper drives:
reset(seed=...) -> (obs, info); step(a) -> (obs, reward, terminated, truncated, info)) used to run the UNMODIFIED
reference trainer end to end (PPO.__init__ -> rollout -> learn) for the boundary tests and goldens."""
import numpy as np


def make_toy_env_class(Box, Discrete):
    class ToyEnv:
        """obs = 4 floats driven by a private numpy generator; the episode terminates when |x0| > 1.2 or after
        `horizon` steps (truncation); reward = 1 - |x0|.  Discrete(2) pushes x0 left / right (CartPole-like)."""
        metadata = {"render_modes": []}
        render_mode = None
        spec = None

        def __init__(self, horizon=40, continuous=False):
            self.observation_space = Box(-np.inf, np.inf, (4,), np.float32)
            self.action_space = Box(-1.0, 1.0, (2,), np.float32) if continuous else Discrete(2)
            self.continuous = continuous
            self.horizon = horizon
            self.rng = np.random.default_rng(0)
            self.state = np.zeros(4, np.float32)
            self.t = 0

        def reset(self, seed=None, options=None):
            if seed is not None:
                self.rng = np.random.default_rng(seed)
            self.state = (self.rng.standard_normal(4) * 0.1).astype(np.float32)
            self.t = 0
            return self.state.copy(), {}

        def step(self, action):
            if self.continuous:
                push = float(np.asarray(action).reshape(-1)[0]) * 0.1
            else:
                push = 0.1 if int(np.asarray(action).reshape(-1)[0]) == 1 else -0.1
            noise = (self.rng.standard_normal(4) * 0.05).astype(np.float32)
            self.state = (self.state * np.float32(0.98) + noise).astype(np.float32)
            self.state[0] += np.float32(push)
            self.t += 1
            terminated = bool(abs(self.state[0]) > 1.2)
            truncated = bool(self.t >= self.horizon and not terminated)
            reward = float(1.0 - abs(self.state[0]))
            return self.state.copy(), reward, terminated, truncated, {}

        def render(self):
            return None

        def close(self):
            pass

    return ToyEnv
