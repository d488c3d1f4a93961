"""Drive the UNMODIFIED reference below its PettingZoo wrapper.

Test infrastructure, build container only.  `RefEnv` seeds the reference's global RNGs,
constructs `CookingEnvironment`, and exposes the same small surface as
oracle/cz_oracle.OracleEnv so lockstep tests can treat both alike.
"""
import random

import numpy as np

from . import ref_dump
from .ref_loader import load_reference


class RefEnv:
    def __init__(self, seed, level, meta_file, num_agents, max_steps, recipes,
                 end_condition_all_dishes=False, action_scheme="scheme3", reward_scheme=None,
                 agent_respawn_rate=0.0, grace_period=20, agent_despawn_rate=0.0):
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

    def layout(self):
        return ref_dump.describe_layout(self.env)

    def step(self, actions):
        """actions: one per agent slot; the reference takes only the active agents' entries."""
        act = [int(a) for i, a in enumerate(actions) if self.env.world.active_agents[i]]
        self.env.accumulated_step(act)
        return ref_dump.step_outputs(self.env)

    def observe_all(self):
        return ref_dump.observe_all(self.env)

    def export_state(self):
        return ref_dump.dump_state(self.env)

    def teleport(self, i, x, y):
        self.env.world.agents[i].move_to((x, y))
