"""The bench JSON line contract, checked on the committed records of the round (profiles/r01_bench_*.json) - CPU only."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RECORDS = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r01_bench_n[0-9].json')))


@pytest.mark.parametrize('path', RECORDS, ids=[os.path.basename(p) for p in RECORDS])
def test_committed_bench_record_keeps_the_contract(path):
    d = json.loads(open(path).read().strip().splitlines()[-1])
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks'):
        assert key in d, key
    assert d['unit'] == 'transitions/s' and d['higher_is_better'] is True and d['scaling'] == 'weak'
    assert d['vs_baseline'] is None                        # BASELINE.md publishes no number for this metric
    assert d['data'] == 'synthetic' and d['dtype'] == 'f32' and 'workload' in d['config'] and 'model' not in d['config']
    assert d['warmup'] >= 3 and d['gpu_launches'] >= d['steps'] > 0
    assert d['n_gpus'] == int(os.path.basename(path)[len('r01_bench_n')])
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert 0.5 < r['frac'] < 1.0 and (r['traffic'] is None or r['traffic'] > 0)
    # value is the whole-job aggregate: rows per step over the max-over-ranks step time
    rows = r['algorithmic_bytes'] / 828 if 'algorithmic_bytes' in r else None
    if rows:
        assert abs(d['value'] - d['n_gpus'] * rows / (d['ms_per_step'] * 1e-3)) / d['value'] < 1e-6
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert 0 < e['value'] < d['value']                     # end to end through the plugin API is never the kernel-only number
    c = d['clocks']
    assert c['sm_mhz'] and c['sm_max_mhz'] and not set(c['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown',
                                                                       'sw_thermal_slowdown'}
    assert c['sm_mhz'] >= 0.9 * c['sm_max_mhz']
    if d['n_gpus'] == 1:
        b = d['cpu_baseline']
        assert b['kind'] == 'port' and b['cores'] >= 1 and b['unit'] == d['unit'] and b['sample']
    assert isinstance(base, dict)


def test_reference_arm_record_keeps_the_contract():
    path = os.path.join(ROOT, 'profiles', 'r01_bench_reference_arm.json')
    if not os.path.exists(path):
        pytest.skip('no reference-arm record committed')
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert d['impl'] == 'reference' and d['unit'] == 'transitions/s' and d['higher_is_better'] is True
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['value'] == d['value']
