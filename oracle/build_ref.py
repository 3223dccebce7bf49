"""TEST / BASELINE INFRASTRUCTURE ONLY -- make the UNMODIFIED reference importable on the GPU box.

The reference is pure Python (no compiled code), so its "build" is a copy: when /root/reference is present (this
container; the GPU box only uses what was copied here) the reference's own source files for the path are copied,
byte for byte, into oracle/_ref/ -- git-ignored, so they never enter this repository's history, but not
gpurun-ignored, so they travel to the GPU box like a built .so.  `bench.py --impl reference` and `cpu_baseline` then
time the reference's OWN modules (architecture.py, principal_neighbourhood_aggregation.py, dqn_agent.py, the
environment) on the host cores, with oracle/shim/ standing in for the three absent third-party packages
(torch_geometric, torch_scatter, gym).  Nothing in the product path imports from here.

    python -m oracle.build_ref
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('GCRL_REFERENCE_ROOT', '/root/reference')
REF_DST = os.path.join(HERE, '_ref')
FILES = ['architecture.py', 'principal_neighbourhood_aggregation.py', 'dqn_agent.py', 'utils.py',
         'test_policy_functions.py', 'dataset_functions.py', 'construct_leighton_graph.py', 'construct_queen_graph.py',
         'construct_mymethod_graph.py', 'construct_HC_dsatur_graph.py']
PACKAGE = 'gym-graph-colouring'


def build_ref():
    """Returns the path of oracle/_ref when it holds the reference's files (freshly copied, or already there)."""
    if os.path.isfile(os.path.join(REF_SRC, 'architecture.py')) and os.path.abspath(REF_SRC) != os.path.abspath(REF_DST):
        os.makedirs(REF_DST, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DST, f))
        dst_pkg = os.path.join(REF_DST, PACKAGE)
        if os.path.isdir(dst_pkg):
            shutil.rmtree(dst_pkg)
        shutil.copytree(os.path.join(REF_SRC, PACKAGE), dst_pkg, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    return REF_DST if os.path.isfile(os.path.join(REF_DST, 'architecture.py')) else None


if __name__ == '__main__':
    print(build_ref())
