"""CPU checks of the measurement harness: the reference arm of bench.py (the oracle timed on the host cores) emits the
contract's JSON line, and the inference oracle agrees with scipy / torch on its own small cases."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_emits_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--size', '64', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith('{')][-1])
    assert line['impl'] == 'reference' and line['metric'].startswith('train images/sec') and line['unit'] == 'images/s'
    assert line['value'] > 0 and line['higher_is_better'] is True and line['gpu_launches'] == 0
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and 'sample' in cb
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--size', '64',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_inference_oracle_small_cases():
    sys.path.insert(0, ROOT)
    from oracle import infer as oi
    # padding forced negative, heights below the minimum zeroed (inferencing/adaptive_scaling.py:145-169)
    logit = torch.tensor([[4.0, -4.0, 4.0], [4.0, 4.0, 4.0]])
    height = torch.tensor([[5.0, 5.0, 2.0], [5.0, 5.0, 5.0]])
    mask, hmap, shape = oi.rough_postprocess(logit, height, image_height=2, image_width=4, padded_height=4, padded_width=6)
    assert shape == (1, 2)
    assert mask.tolist() == [[1, 0, 0], [0, 0, 0]] and hmap.tolist() == [[5.0, 5.0, 0.0], [0.0, 0.0, 0.0]]
    # peaks: plateaus count, sub-threshold maxima do not (inferencing/adaptive_scaling.py:477-491)
    score = np.zeros((7, 7), dtype=np.float32)
    score[1, 1] = 0.9
    score[1, 2] = 0.9
    score[5, 5] = 0.6
    assert oi.peak_mask(score, None, size=5, positive_thr=0.7).nonzero()[0].tolist() == [1, 1]
    cm = np.ones((7, 7), dtype=np.uint8)
    cm[1, 1] = 0
    assert list(zip(*oi.peak_mask(score, cm, size=5, positive_thr=0.7).nonzero())) == [(1, 2)]
    # softmax / sigmoid / permutes of the precise pass (:322-386)
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(1, c, 4, 6, generator=g) for c in (1, 2, 4, 4)]
    prob, offs, angs, dists = oi.precise_postprocess(*feats, image_height=7, image_width=12, padded_height=32, padded_width=32)
    assert prob.shape == (4, 6) and offs.shape == (4, 6, 2) and angs.shape == (4, 6, 4) and dists.shape == (4, 6, 4)
    assert np.allclose(angs.sum(-1), 1.0, atol=1e-6)
    assert np.allclose(prob, torch.sigmoid(feats[0][0, 0]).numpy())


def test_roofline_of_picks_the_largest_call_and_divides_algorithmic_work_by_measured_time():
    """bench.roofline_of on a synthetic per-call table: the dominant call decides the bound (HBM bytes for a streaming
    kernel, tensor FLOPs for a tcgen05 GEMM), `achieved` = algorithmic work per launch / average launch duration, shares are
    taken against the (bracketed) step the per-call times were measured in."""
    sys.path.insert(0, ROOT)
    import bench
    peaks = {'hbm_gbs': 6000.0, 'tflops_burst': 1600.0, 'tflops_sustained': 1400.0, 'source': 'test'}
    steps, step_ms = 2, 50.0
    table = {
        'stream A': {'entry': 'vkocr_head_combine_fwd', 'calls': 2, 'ms': 20.0, 'flops': 0.0, 'bytes': 2 * 15e9},
        'gemm B': {'entry': 'vkocr_gemm_nt', 'calls': 4, 'ms': 12.0, 'flops': 4 * 3e12, 'bytes': 0.0},
        'tiny': {'entry': 'vkocr_zero', 'calls': 10, 'ms': 0.1, 'flops': 0.0, 'bytes': 0.0},
    }
    r = bench.roofline_of(table, steps, step_ms, peaks)
    assert r['bound'] == 'hbm' and r['kernel'].startswith('vkocr_head_combine_fwd') and r['unit'] == 'GB/s'
    assert abs(r['launch_ms'] - 10.0) < 1e-9 and abs(r['achieved'] - 1500.0) < 1e-6 and abs(r['frac'] - 0.25) < 1e-9
    assert abs(r['share_of_step'] - 0.2) < 1e-9 and abs(r['all_gemm_share_of_step'] - 0.12) < 1e-9
    assert abs(r['all_gemm_tflops'] - 1000.0) < 1e-6 and r['algorithmic_bytes'] == 15e9
    table['gemm B']['ms'] = 40.0
    r = bench.roofline_of(table, steps, step_ms, peaks)
    assert r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s' and abs(r['achieved'] - 300.0) < 1e-6
    assert abs(r['frac'] - 300.0 / 1400.0) < 1e-9 and r['algorithmic_flops'] == 3e12
    assert bench.roofline_of({'tiny': table['tiny']}, steps, step_ms, peaks) is None
