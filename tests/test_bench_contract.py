"""CPU check of the bench.py output contract on the reference arm (the native arm needs a GPU): one JSON line on stdout
with the keys the driver reads."""
import json
import os
import subprocess
import sys

from tests.util import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS='4')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['unit'] == 'images/s' and line['higher_is_better'] is True
    assert line['metric'].startswith('images/sec localized') and line['value'] > 0
    assert line['steps'] == 1 and line['warmup'] == 0 and line['n_gpus'] == 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    base = line['cpu_baseline']
    assert base['kind'] == 'port' and base['cores'] >= 1 and base['value'] == line['value'] and 'frames' in base['sample']
    assert line['config']['workload'] == 'batch32_480x720_forward+dsac256'


def test_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like the native one (torchrun, N ranks): rank 0 alone works and prints."""
    env = dict(os.environ, OMP_NUM_THREADS='2')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29631', os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
           '--warmup', '0']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['n_gpus'] == 2 and line['value'] > 0


def test_both_arms_describe_the_same_config():
    """`config` is built by one function for both arms, so the driver's same-config check holds by construction."""
    import argparse
    import bench
    args = argparse.Namespace(batch=32, hyps=256)
    cfg = bench.config_of(args, 1)
    assert cfg['workload'] == 'batch32_480x720_forward+dsac256' and 'l2' in cfg and 'parallelism' in cfg
    src = open(os.path.join(ROOT, 'bench.py')).read()
    assert src.count("'config': config_of(args,") == 2   # the native and the reference line
