"""GPU test of the drop-in entry point: `python src/main.py` with the reference's flags runs one complete
communication round (server epoch, extraction, one image / one text / one multimodal client with inter+intra contrast,
con_w aggregation, distillation, COCO-1K-style evaluation) on small synthetic data, entirely on the CUDA path."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_main_runs_one_round(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    cmd = [sys.executable, str(ROOT / 'src' / 'main.py'), '--name', str(tmp_path / 'smoke'), '--server_lr', '1e-5', '--seed', '0',
           '--feature_dim', '256', '--pub_data_num', '256', '--agg_method', 'con_w', '--contrast_local_intra',
           '--contrast_local_inter', '--num_img_clients', '1', '--num_txt_clients', '1', '--num_mm_clients', '1',
           '--client_num_per_round', '3', '--local_epochs', '1', '--comm_rounds', '1', '--interintra_weight', '0.5',
           '--kd_weight', '0.3', '--private_samples', '3000', '--image_size', '64', '--client_image_size', '64',
           '--test_images', '200', '--test_folds', '2', '--pub_batch_size', '64']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'start distilling' in res.stdout and '[Eval] Report @step 1' in res.stdout
    assert 'recall_1' in res.stdout


def test_conw_weights_are_uniform_for_identical_clients_full_round_shape():
    """Round-level invariant at the full public-set size: identical client representations -> aggregate == input."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200 import ops
    g = torch.Generator().manual_seed(0)
    v = torch.nn.functional.normalize(torch.randn(50000, 256, generator=g), dim=1).cuda()
    gl = torch.nn.functional.normalize(torch.randn(50000, 256, generator=g), dim=1).cuda()
    out = ops.conw_aggregate([v, v, v], gl)
    assert torch.allclose(out, v, atol=1e-6)
