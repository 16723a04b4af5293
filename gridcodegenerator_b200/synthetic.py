"""Seeded synthetic states: q ~ U(-pi,pi), qd ~ U(-2,2),
u ~ U(-20,20), qdd ~ U(-5,5), drawn in float64 and rounded to float32."""
import numpy as np

CONFIG_SEED_BASE = 20261018
CONFIG_INDEX = {"iiwa14": 1, "hyq": 2, "atlas": 3, "chain64": 4, "mixed5": 5, "pchain4": 6}


def make_states(n: int, num_states: int, seed: int):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, (num_states, n)).astype(np.float32)
    qd = rng.uniform(-2.0, 2.0, (num_states, n)).astype(np.float32)
    u = rng.uniform(-20.0, 20.0, (num_states, n)).astype(np.float32)
    qdd = rng.uniform(-5.0, 5.0, (num_states, n)).astype(np.float32)
    return q, qd, u, qdd


def seed_for(robot_name: str) -> int:
    return CONFIG_SEED_BASE + CONFIG_INDEX.get(robot_name, 0)


def pack_q_qd_u(q, qd, u):
    """State-major [q | qd | u], stride 3n (reference gridData::d_q_qd_u)."""
    return np.ascontiguousarray(np.concatenate([q, qd, u], axis=1), dtype=np.float32)


def pack_q_qd(q, qd):
    return np.ascontiguousarray(np.concatenate([q, qd], axis=1), dtype=np.float32)
