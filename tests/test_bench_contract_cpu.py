"""The driver's reference arm (`bench.py --impl reference`) is CPU-only, so
its JSON contract is checked here: one line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    result = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
         '--steps', '1', '--warmup', '1', '--utterances', '6'],
        cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stderr[-2000:]
    lines = [line for line in result.stdout.splitlines() if line.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference'
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
                'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
                'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['metric'] == 'audio-sec/sec' and line['unit'] == 'audio-s/s'
    assert line['value'] > 0 and line['higher_is_better'] is True
    assert line['cpu_baseline']['kind'] in ('port', 'reference')
    assert line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['sample']
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['e2e']['value'] == line['value']
    assert 'workload' in line['config']
