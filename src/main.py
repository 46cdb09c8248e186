"""Entry point with the reference's command line (src/main.py:36-134): `python src/main.py --name ... --contrast_local_intra
--contrast_local_inter ...`.  Every flag of the reference is accepted; the ones its code path never reads
 are parsed and ignored here too.  Extra flags size the synthetic data (no dataset is on the box)."""
import argparse
import os
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'src'))

parser = argparse.ArgumentParser(description='Federated Learning')


def args():
    p = parser.add_argument
    p('--name', type=str, default='Test')
    p('--exp_dir', type=str, default='./experiments/')
    p('--local_epochs', type=int, default=5)
    p('--comm_rounds', type=int, default=30)
    p('--model', type=str, default='resnet34')
    p('--img_model_local', type=str, default='resnet10')
    p('--pretrained', type=int, default=0)
    p('--no-cuda', action='store_true', default=False)
    p('--seed', type=int, default=random.randint(0, 100000))
    p('--device', type=int, default=0)
    p('--num_img_clients', type=int, default=10)
    p('--num_txt_clients', type=int, default=10)
    p('--num_mm_clients', type=int, default=15)
    p('--client_num_per_round', type=int, default=10)
    p('--dataset', type=str, default='cifar100', choices=['svhn', 'cifar10', 'cifar100'])
    p('--data_root', type=str, default=os.environ.get('HOME', '.') + '/data/')
    p('--batch_size', type=int, default=64)
    p('--alpha', type=float, default=0.5)
    p('--server_lr', type=float, default=0.0002)
    p('--lr', type=float, default=0.1)
    p('--loss', type=str, default='l1', choices=['l1', 'kl', 'l1softmax'])
    p('--scheduler', type=str, default='multistep', choices=['multistep', 'cosine', 'exponential', 'none'])
    p('--steps', nargs='+', default=[0.05, 0.15, 0.3, 0.5, 0.75], type=float)
    p('--scale', type=float, default=0.1)
    p('--weight_decay', type=float, default=5e-4)
    p('--momentum', type=float, default=0.9)
    p('--log_interval', type=int, default=10)
    p('--save_interval', type=int, default=10)
    p('--disable_distill', action='store_true', default=False)
    p('--agg_method', type=str, default='con_w')
    p('--contrast_local_intra', action='store_true', default=False)
    p('--contrast_local_inter', action='store_true', default=False)
    p('--mlp_local', action='store_true', default=False)
    p('--kd_weight', type=float, default=0.3)
    p('--interintra_weight', type=float, default=0.5)
    p('--loss_scale', action='store_true', default=False)
    p('--save_client', action='store_true', default=False)
    p('--data_local', action='store_true', default=False)
    p('--pub_data_num', type=int, default=50000)
    p('--feature_dim', type=int, default=256)
    p('--not_bert', action='store_true', default=False)
    # synthetic-data sizing (not in the reference)
    p('--private_samples', type=int, default=50000, help='synthetic private pool partitioned over the clients')
    p('--image_size', type=int, default=224)
    p('--client_image_size', type=int, default=256, help='CIFAR clients upsample to 256 (load_FL_datasets.py:16-21)')
    p('--test_images', type=int, default=5000)
    p('--test_folds', type=int, default=5)
    p('--pub_batch_size', type=int, default=128)
    p('--no_cuda_graphs', action='store_true', default=False,
      help='launch every kernel eagerly (default: each step function is captured as a CUDA graph, the path bench.py times)')


args()

if __name__ == '__main__':
    a = parser.parse_args()
    import numpy as np
    import torch
    from algorithms.MMFL import MMFL

    random.seed(a.seed)
    np.random.seed(a.seed)
    torch.manual_seed(a.seed)                                  # helper.set_seed (utils/helper.py:136-144)
    Algo = MMFL(a, None)
    Algo.create_model(a)
    Algo.load_dataset(a)
    for round_n in range(a.comm_rounds):
        Algo.train(round_n)
    Algo.logger.log('Best:')
    Algo.engine.report_scores(step=a.comm_rounds, scores=Algo.best_scores, metadata=Algo.best_metadata)
