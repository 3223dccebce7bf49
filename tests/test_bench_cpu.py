"""CPU test of bench.py's output contract through the reference arm (the only arm that runs without a GPU): exactly
ONE line on stdout, JSON, with the keys the driver reads; whatever libraries print goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--vertices', '24', '--cpu-sample-graphs', '2'], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'qnet_fwd_bwd_graphs_per_sec' and d['unit'] == 'graphs/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'graphs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_non_zero_ranks_of_the_reference_arm_exit_silently():
    e = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                       cwd=ROOT, env=e, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''
