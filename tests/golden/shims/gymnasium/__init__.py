"""gymnasium.make stand-in over oracle/cartpole_env.py (TEST INFRASTRUCTURE, see ../pennylane.py)."""
from oracle.cartpole_env import CartPoleV1


def make(env_name, render_mode=None):
    assert env_name == "CartPole-v1"
    return CartPoleV1()
