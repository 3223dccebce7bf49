"""graph_colouring_with_rl_b200 -- the GN Q-network hot path of
gpdwatkins/graph_colouring_with_RL (ReLCol) as hand-written sm_100a CUDA kernels behind the
reference's own Python API.

    from graph_colouring_with_rl_b200.architecture import GNNetwork, GNNBlock
    from graph_colouring_with_rl_b200.principal_neighbourhood_aggregation import AGGREGATORS, SCALERS
    from graph_colouring_with_rl_b200.dqn_agent import DQNAgentGN, ReplayBuffer
    from graph_colouring_with_rl_b200.envs import GraphColouring, BatchedColouringEnv, make

Submodules import torch and the CUDA library lazily; importing the package itself is free.
There is no CPU fallback: every compute entry point raises without a CUDA device.
"""
__version__ = '0.1.0'
