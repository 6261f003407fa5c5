"""The drop-in boundary end to end: the reference's OWN driver scripts (vendored unmodified into oracle/_ref by
tools/vendor_ref.py) run on top of this repository's `tiger` package - `load_jodie_data` on a toy dataset in the
on-disk format, `init_data` / `init_model`, the training loop (`loss.backward()` = the native step as one autograd
node, torch.optim.Adam), `flush_msg`, `save/load_memory_state`, `eval_edge_prediction`, checkpoints and results."""
import glob
import json
import os
import subprocess
import sys

import pytest
import torch

from jodie_utils import write_toy_dataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')
MIRROR = os.path.join(ROOT, 'www2023tiger_b200')

LAUNCH = r'''
import runpy, sys
mirror, ref, script = sys.argv[1:4]
# `tiger` resolves to this repository's package, `init_utils` / `train_utils` / `CHANGELOG` to the reference's files
sys.path[:0] = [mirror, ref, sys.argv[4]]
sys.argv = [script] + sys.argv[5:]
import tiger
assert tiger.__file__.startswith(mirror), tiger.__file__
runpy.run_path(script, run_name='__main__')
'''


def run_driver(script, tmp_path, *args, timeout=900):
    cmd = [sys.executable, '-c', LAUNCH, MIRROR, REF, os.path.join(REF, script), ROOT, *args]
    return subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=timeout)


def check_results(tmp_path, n_epochs):
    files = glob.glob(os.path.join(str(tmp_path), 'results', '*.json'))
    assert len(files) == 1, files
    res = json.load(open(files[0]))
    for key in ('val_ap', 'test_ap', 'ind_val_ap', 'ind_test_ap'):
        assert key in res and 0.0 <= res[key] <= 1.0, (key, res)
    assert res['test_ap'] > 0.5                       # better than chance after a little training
    assert glob.glob(os.path.join(str(tmp_path), 'saved_models', '*.pth'))
    return res


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tiger')), reason='oracle/_ref missing (tools/vendor_ref.py)')
@pytest.mark.parametrize('restarter', ['seq', 'static'])
def test_reference_training_script_runs_on_the_drop_in(tmp_path, restarter):
    write_toy_dataset(str(tmp_path), n_events=4000)
    res = run_driver('train_self_supervised.py', tmp_path, '-d', 'toy', '--root', str(tmp_path), '--n_epochs', '2',
                     '--bs', '100', '--hist_len', '8', '--n_neighbors', '5', '--restarter_type', restarter,
                     '--restart_prob', '0.05', '--seed', '0')
    assert res.returncode == 0, res.stderr[-3000:]
    out = check_results(tmp_path, 2)
    # the checkpoint the driver saved loads into the reference's key layout (state_dict compatibility)
    sd = torch.load(glob.glob(os.path.join(str(tmp_path), 'saved_models', '*.pth'))[0], map_location='cpu')
    assert 'left_memory.vals' in sd and 'right_mem_updater.cell.weight_ih' in sd and 'score_fn.fc1.weight' in sd
    assert any(k.startswith('restarter_fn.') for k in sd)
    assert out['test_auc'] >= 0.0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='the DDP driver needs two GPUs')
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tiger')), reason='oracle/_ref missing (tools/vendor_ref.py)')
def test_reference_ddp_training_script_runs_on_the_drop_in(tmp_path):
    write_toy_dataset(str(tmp_path), n_events=6000)
    res = run_driver('train_self_supervised_ddp.py', tmp_path, '-d', 'toy', '--root', str(tmp_path), '--n_epochs', '2',
                     '--bs', '100', '--hist_len', '8', '--n_neighbors', '5', '--gpu', '0,1', '--port', '29533',
                     '--seed', '0')
    assert res.returncode == 0, res.stderr[-3000:]
    check_results(tmp_path, 2)
