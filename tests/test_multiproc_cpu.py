"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the task dealing, the gather of the
per-task results and the max-over-ranks timing that bench.py and the model-selection batch use."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import numpy as np
from signer_b200 import dist
rank, world, local = dist.init(backend="gloo")
tasks = [dict(K=k, burn=10, eval=10, seed=100 + k) for k in range(2, 12)]
def runner(ts):                      # stands in for signer_gibbs_batch on this rank's GPU
    return [np.full(3, float(t["K"]) * 10 + rank) for t in ts]
res = dist.run_tasks_distributed(tasks, runner)
mx = dist.max_over_ranks(1.5 + rank)
dist.barrier()
if rank == 0:
    print(json.dumps(dict(world=world, ks=[float(r[0]) // 10 for r in res], ranks=[int(r[0]) %% 10 for r in res], mx=mx)))
"""


def test_deal_tasks_is_lpt_and_balanced():
    from signer_b200 import dist
    costs = [k * 2000 for k in range(2, 41)] * 4          # config 5: nsig 2..40 x 4 chains
    bins = dist.deal_tasks(costs, 8)
    assert sorted(t for b in bins for t in b) == list(range(len(costs)))
    load = [sum(costs[t] for t in b) for b in bins]
    assert max(load) / (sum(load) / 8) < 1.02             # >= 7.8x of 8 on paper
    assert dist.deal_tasks([5, 1, 1, 1], 2) == [[0], [1, 2, 3]]
    assert dist.deal_tasks([], 3) == [[], [], []]


def test_two_rank_gloo_run(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["ks"] == [float(k) for k in range(2, 12)]
    assert set(d["ranks"]) == {0, 1}                      # both ranks did work, results back in task order
    assert d["mx"] == 2.5
