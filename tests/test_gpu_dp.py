"""Data-parallel training on real GPUs: two ranks over NCCL, bucketed gradient all-reduce started
from the backward hooks, captured together with the SGD step in the training CUDA graph.  Needs two GPUs (skipped
otherwise; the host-side logic is covered on the CPU over gloo by tests/test_cpu_host.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

WORKER = r"""
import os, sys, json, copy, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from wesup_b200 import parallel, synth
from wesup_b200.models import initialize_trainer
rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.backends.cudnn.allow_tf32 = False
H, W = 160, 192
data = [tuple(t.to(dev) for t in synth.sample(H, W, index=10 * rank + i, ratio=2e-3)) for i in range(3)]

def make(graph, overlap):
    torch.manual_seed(0)
    t = initialize_trainer("wesup", device=dev, pretrained=False, materialize_hypercolumn=False, cuda_graph=graph, metrics_lag=0)
    t.optimizer, _ = t.get_default_optimizer()
    t.enable_data_parallel(overlap=overlap)
    return t

def params(t):
    return torch.cat([p.detach().flatten() for p in t.model.parameters()])

ok, notes = True, []
# (1) one eager iteration: averaged gradient == mean over ranks of the local gradients
t = make(False, True)
ref = copy.deepcopy(t.model)
(x, sp), (pm, sl) = t.preprocess(*data[0])
pred = ref((x, sp)); t.model.sp_features, t.model.sp_pred = ref.sp_features, ref.sp_pred
loss = t.compute_loss(pred, (pm, sl)); loss.backward()
local_g = torch.cat([p.grad.flatten() for p in ref.parameters()])
mean_g = local_g.clone(); dist.all_reduce(mean_g); mean_g /= world
t.train_one_iteration("train", *data[0])
got = torch.cat([p.grad.flatten() for p in t.model.parameters()])
err = float((got - mean_g).abs().max() / (mean_g.abs().max() + 1e-30))
ok &= err < 1e-5; notes.append(("grad_vs_mean", err))
ok &= len(t.grad_sync.buckets) >= 3; notes.append(("buckets", len(t.grad_sync.buckets)))
# (2) graph path vs eager path vs blocking path: same parameters after 7 iterations, identical on all ranks
finals = {{}}
for name, graph, overlap in (("graph_overlap", True, True), ("eager_overlap", False, True), ("eager_blocking", False, False)):
    t = make(graph, overlap)
    for i in range(7):
        t.train_one_iteration("train", *data[i % 3])
    t.flush_metrics()
    torch.cuda.synchronize()
    p = params(t)
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok &= same; notes.append((name + "_ranks_identical", same))
    finals[name] = p
    if graph:
        n_graphs = len(getattr(t, "_graphs", {{}}))
        ok &= n_graphs >= 1; notes.append(("graphs_captured", n_graphs))
    losses = t.tracker.history["loss"]
    ok &= all(l == l for l in losses)
for a, b in (("graph_overlap", "eager_overlap"), ("eager_overlap", "eager_blocking")):
    err = float((finals[a] - finals[b]).abs().max() / finals[b].abs().max())
    ok &= err < 1e-5; notes.append((a + "_vs_" + b, err))
print("RANK %d %s %s\n" % (rank, "OK" if ok else "FAIL", json.dumps(notes)), end="", flush=True)
del t, finals
import gc; gc.collect(); torch.cuda.synchronize()
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_training_on_two_gpus(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    env.pop("NCCL_DEBUG", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "RANK 0 OK" in r.stdout and "RANK 1 OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
