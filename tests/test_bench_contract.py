"""The reference arm of bench.py runs on CPU: its JSON line must carry the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.strip().splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d['impl'] == 'reference' and d['unit'] == 'tokens/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('shape-program tokens/sec') and d['value'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    staged = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'plankassembly', 'models.py'))
    assert cb['kind'] == ('reference' if staged else 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert 'cpu' in cb and d['ms_per_step'] > 0 and d['steps'] == 1
    # same config block as the b200 arm prints (the driver compares them)
    assert 'BASELINE configs[1]' in d['config']['workload'] and d['config']['global_batch'] == 64
