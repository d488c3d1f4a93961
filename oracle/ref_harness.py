"""Drive the UNMODIFIED reference below its PettingZoo wrapper.

Test infrastructure, build container only.  `RefEnv` seeds the reference's global RNGs,
constructs `CookingEnvironment`, and exposes the same small surface as
oracle/cz_oracle.OracleEnv so lockstep tests can treat both alike.
"""
import contextlib
import random

import numpy as np

from . import ref_dump
from .ref_loader import load_reference


class RefEnv:
    def __init__(self, seed, level, meta_file, num_agents, max_steps, recipes,
                 end_condition_all_dishes=False, action_scheme="scheme3", reward_scheme=None,
                 agent_respawn_rate=0.0, grace_period=20, agent_despawn_rate=0.0, spawn_stream=None):
        ce = load_reference()
        random.seed(seed)
        np.random.seed(seed)
        self.env = ce.CookingEnvironment(
            level=level, meta_file=meta_file, num_agents=num_agents, max_steps=max_steps,
            recipes=list(recipes), obs_spaces=["feature_vector"] * num_agents,
            end_condition_all_dishes=end_condition_all_dishes, action_scheme=action_scheme,
            reward_scheme=reward_scheme, agent_respawn_rate=agent_respawn_rate,
            grace_period=grace_period, agent_despawn_rate=agent_despawn_rate)
        self.env.reset()
        self.num_agents = num_agents
        self.spawn_stream = spawn_stream

    def layout(self):
        return ref_dump.describe_layout(self.env)

    def step(self, actions):
        """actions: one per agent slot; the reference takes only the active agents' entries."""
        act = [int(a) for i, a in enumerate(actions) if self.env.world.active_agents[i]]
        with self._patched_rng():
            self.env.accumulated_step(act)
        return ref_dump.step_outputs(self.env)

    @contextlib.contextmanager
    def _patched_rng(self):
        """Route the two RNG call sites of handle_agent_spawn (np.random.random at
        cooking_world.py:274,276 and random.sample at engine/parsing.py:157-158) to the shared
        counter-based stream for the duration of one step.  No reference code is modified."""
        if self.spawn_stream is None:
            yield
            return
        from cooking_zoo.cooking_world.engine import parsing
        stream = self.spawn_stream
        stream.begin_step(self.env.t + 1)

        class _Random:
            @staticmethod
            def sample(seq, k):
                assert k == 1
                return [stream.choice(seq)]

            @staticmethod
            def random():
                return stream.uniform()

        saved_np, saved_mod = np.random.random, parsing.random
        np.random.random, parsing.random = stream.uniform, _Random
        try:
            yield
        finally:
            np.random.random, parsing.random = saved_np, saved_mod

    def observe_all(self):
        return ref_dump.observe_all(self.env)

    def export_state(self):
        return ref_dump.dump_state(self.env)

    def teleport(self, i, x, y):
        self.env.world.agents[i].move_to((x, y))
