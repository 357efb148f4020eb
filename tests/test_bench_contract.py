"""The JSON lines bench.py prints carry the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup',
             'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
             'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}
E2E_KEYS = {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}


def run_bench(*args, timeout=900):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')]
                         + list(args), capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line():
    # C2 frames (2^20 points x 16 series) keep the CPU test short; the default
    # workload (C4, 2^24-point frames) runs the same code.
    line = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0',
                     '--workload', 'C2')
    assert line['impl'] == 'reference'
    assert BASE_KEYS <= set(line)
    assert E2E_KEYS <= set(line['e2e'])
    assert line['e2e']['h2d_bytes_per_step'] == 0
    assert line['e2e']['value'] == line['value'] > 0
    cpu = line['cpu_baseline']
    assert cpu['kind'] in ('port', 'reference') and cpu['cores'] >= 1
    assert cpu['value'] == line['value'] and cpu['sample']
    assert line['config']['workload'].startswith('C2')


def test_default_workload_is_the_target_config():
    """The driver's plain run measures configs[3] (C4), with identical
    ``config`` objects in both arms."""
    sys.path.insert(0, ROOT)
    import bench
    ap_default = 'C4'
    assert bench.WORKLOADS[ap_default]['log2n'] == 24
    cfg = bench.config_of('C4', bench.WORKLOADS['C4'])
    assert cfg['workload'].startswith('C4') and cfg['fft_length'] == 1 << 24
    import inspect
    assert "default='C4'" in inspect.getsource(bench.main)


@pytest.mark.gpu
def test_b200_arm_line():
    line = run_bench('--steps', '2', '--warmup', '3', '--no-cpu')
    assert BASE_KEYS | {'gpu_launches', 'clocks', 'roofline'} <= set(line)
    assert E2E_KEYS <= set(line['e2e'])
    assert line['n_gpus'] == 1 and line['steps'] == 2 and line['warmup'] >= 3
    assert line['value'] > line['e2e']['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] > 0
    assert line['gpu_launches'] >= 4 * line['steps']
    roof = line['roofline']
    assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(roof)
    assert roof['bound'] == 'hbm' and 0 < roof['frac'] < 1
    assert abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-9
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(line['clocks'])
    assert line['config']['workload'].startswith('C4')
    assert line['strong']['value'] > 0
