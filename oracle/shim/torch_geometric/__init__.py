"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for `torch_geometric` (not installed here).
Restates exactly the names the reference imports (architecture.py:5-7, dqn_agent.py:5,
graph_colouring_env.py:6, utils.py:12).  "parity unpinned" at this boundary."""
