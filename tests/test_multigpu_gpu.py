"""Haplotype sharding over 2 GPUs with NCCL: scatter the int8 block, run the hot path per
rank with no collective, gather labels -- identical to the single-GPU result.  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from gnomix_b200 import parallel, GBTForest
from tests import util
rng = np.random.default_rng(21)
C, M, A, S, N = 12007, 300, 7, 9, 333
coefs, icpts, ctx = util.random_lr(rng, C, M, A)
base = util.make_lr_base(C, M, A, coefs, icpts)
from gnomix_b200.smooth import XGB_Smoother
sm = XGB_Smoother(n_windows=C // M, num_ancestry=A, smooth_window_size=S)
sm.model = GBTForest.random(rng, A, S, n_rounds=20, depth=4)
X = torch.from_numpy(util.random_haplotypes(rng, N, C)).cuda()
mine = parallel.scatter_rows(X if rank == 0 else None, N, C, torch.int8, "cuda")
lab = sm.predict(base.predict_proba(mine))
full = parallel.gather_rows(lab, N)
if rank == 0:
    want = sm.predict(base.predict_proba(X))
    assert torch.equal(full, want)
    print("OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gpu_scatter_predict_gather(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK" in out.stdout
