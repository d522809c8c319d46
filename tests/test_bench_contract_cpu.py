"""CPU: `bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's keys,
and the GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

from conftest import REPO


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), *args], capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_contract_line():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'stroke_patches_per_sec_128x128' and d['unit'] == 'patches/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = _run('--steps', '1', '--warmup', '0')
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
