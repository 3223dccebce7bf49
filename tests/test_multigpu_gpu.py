"""k-GPU gradient == 1-GPU gradient on the union minibatch, on real GPUs over NCCL.
Needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`; skipped on a 1-GPU box.
The checks live in tests/mgpu_worker.py (one process per GPU under torchrun)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('math_mode', ['fp32', 'bf16x3'])
def test_k_rank_gradient_equals_one_rank_gradient(tmp_path, math_mode):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs (gpurun --gpus 2)')
    port = 29600 + os.getpid() % 1000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(HERE, 'mgpu_worker.py'), str(tmp_path), math_mode]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    rep = json.load(open(os.path.join(str(tmp_path), 'mgpu_rank0.json')))
    print(json.dumps(rep))
    out = os.environ.get('GCRL_MGPU_REPORT')
    if out:
        with open(out, 'a') as f:
            f.write(json.dumps(rep) + '\n')


def test_multi_gpu_training_loop_keeps_replicas_identical(tmp_path):
    """dqn_main.DeviceTrainer over 2 GPUs: waves sharded by graph, rank-local replay rings, one global minibatch per learn
    step, cadence from the global step count (tests/mgpu_worker.py train_main)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs (gpurun --gpus 2)')
    port = 29700 + os.getpid() % 1000
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(HERE, 'mgpu_worker.py'), str(tmp_path), 'train']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    rep = json.load(open(os.path.join(str(tmp_path), 'mgpu_train_rank0.json')))
    print(json.dumps(rep))
    assert rep['learn_steps'] > 0 and rep['episodes'] == 48
    out = os.environ.get('GCRL_MGPU_REPORT')
    if out:
        with open(out, 'a') as f:
            f.write(json.dumps(rep) + '\n')
