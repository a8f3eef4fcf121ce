"""bench.py contract on the CPU: the reference arm (the oracle port timed on the host cores, a bounded sample per step) prints
ONE JSON line with the keys the driver reads, on the same metric / unit / workload name as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'genomes/hour' and d['unit'] == 'genomes/hour'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['value'] > 0
    assert d['cpu_baseline']['kind'] in ('port', 'hmmer') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {"value": d['value'], "unit": d['unit'], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert d['config']['workload'] == bench.workload_name(3, bench.total_model_positions(bench.model_db()))
    assert 'no extrapolation' in d['cpu_baseline']['sample'] and d['cpu_baseline']['gcups_per_core'] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
