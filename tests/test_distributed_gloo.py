"""CPU: world_size-2 `gloo` check of the N>1 plumbing bench.py uses (image sharding, max-over-ranks reduction, only
rank 0 reporting).  The head path itself has no collective — this covers the host logic around it."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import bench
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
# every rank owns its own images (weak scaling): batches must differ across ranks and be reproducible per rank
x, gtb, gtl, aid, metas = bench.synth_batch(2, 1234 + rank * 10)
sig = torch.tensor([float(x.sum()), float(gtb[0].sum())])
sigs = [torch.zeros(2) for _ in range(world)]
dist.all_gather(sigs, sig)
t = torch.tensor([10.0 + 5 * rank])          # per-rank elapsed ms -> reported time is the max over ranks
dist.all_reduce(t, op=dist.ReduceOp.MAX)
dist.barrier()
if rank == 0:
    print(json.dumps(dict(distinct=bool((sigs[0] != sigs[1]).any()), tmax=float(t[0]), world=world,
                          value=world * 2 * 1 / (float(t[0]) / 1e3))))
dist.destroy_process_group()
'''


def test_two_rank_gloo(tmp_path):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    w = tmp_path / 'worker.py'
    w.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(w), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = json.loads([l for l in outs[0][0].splitlines() if l.startswith('{')][-1])
    assert line['distinct'] and line['world'] == 2
    assert line['tmax'] == 15.0                      # max over ranks, not rank 0's own 10 ms
    assert abs(line['value'] - 2 * 2 / 0.015) < 1e-6
    assert outs[1][0].strip() == ''                  # only rank 0 prints


GRAD_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pointtinybenchmark_b200.dist import allreduce_grads
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
torch.manual_seed(0)
m = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.Linear(4, 2))
for i, p in enumerate(m.parameters()):
    p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
list(m.parameters())[3].grad = None if rank == 1 else list(m.parameters())[3].grad      # unused on one rank -> zeros there
nbytes = allreduce_grads(m)
vals = [float(p.grad.reshape(-1)[0]) for p in m.parameters()]
same = all(bool((p.grad == p.grad.reshape(-1)[0]).all()) for p in m.parameters())
if rank == 0:
    print(json.dumps(dict(vals=vals, same=same, nbytes=nbytes)))
dist.destroy_process_group()
'''


def test_gradient_allreduce_two_ranks(tmp_path):
    """the head's only collective: one flat-bucket gradient all-reduce (mean over ranks), world_size 2 on gloo"""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    w = tmp_path / 'gworker.py'
    w.write_text(GRAD_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(w), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = json.loads([l for l in outs[0][0].splitlines() if l.startswith('{')][-1])
    # rank r holds (r+1)*(i+1) for parameter i: mean over ranks = 1.5*(i+1); parameter 3 exists on rank 0 only: (4 + 0)/2
    assert line['vals'] == [1.5, 3.0, 4.5, 2.0] and line['same']
    assert line['nbytes'] == (8 * 4 + 4 + 4 * 2 + 2) * 4


BUCKET_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pointtinybenchmark_b200.dist import GradBucket
from pointtinybenchmark_b200.registry import build_head
from pointtinybenchmark_b200 import cpr_head  # noqa
from tests.test_gpu_cpr_head import head_cfg
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
torch.manual_seed(0)
head = build_head(head_cfg(dict(num_classes=4, C=32, stride=8, radius=2)))      # real parameter names: cls_convs.{i}.*, cls_out, ins_out
bucket = GradBucket(head, overlap=(os.environ.get('PTB_TEST_OVERLAP', '1') == '1'))
names = [n for n, p in head.named_parameters()]
ptrs0 = [p.grad.data_ptr() for p in head.parameters()]
out = []
for step in range(2):
    bucket.zero()
    # a fake backward: every parameter gets grad (rank+1)*(index+1) through autograd, except ins_out.* which stays unused on rank 1
    loss = 0
    for i, (n, p) in enumerate(head.named_parameters()):
        if rank == 1 and n.startswith('ins_out'):
            continue
        loss = loss + (p * float((rank + 1) * (i + 1))).sum()
    loss.backward()
    nbytes = bucket.wait()
    out.append([float(p.grad.reshape(-1)[0]) for p in head.parameters()])
views = all(p.grad.data_ptr() == q for p, q in zip(head.parameters(), ptrs0))
if rank == 0:
    print(json.dumps(dict(names=names, vals=out, views=views, nbytes=nbytes, n_buckets=len(bucket.ranges),
                          order_first=[n for n, p in head.named_parameters() if p.grad.data_ptr() == bucket.flat.data_ptr()])))
dist.destroy_process_group()
'''


import pytest  # noqa: E402


@pytest.mark.parametrize('overlap', ['1', '0'])
def test_grad_bucket_overlapped_allreduce_two_ranks(tmp_path, overlap):
    """GradBucket: gradients live as views of one persistent flat buffer, buckets (classifiers, then tower layers last to first) are
    all-reduced from post-accumulate hooks, unused parameters contribute zeros, a second step re-uses the same storage."""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    w = tmp_path / 'bworker.py'
    w.write_text(BUCKET_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), PTB_TEST_OVERLAP=overlap)
        procs.append(subprocess.Popen([sys.executable, str(w), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = json.loads([l for l in outs[0][0].splitlines() if l.startswith('{')][-1])
    assert line['views'] and line['n_buckets'] == 5                       # classifiers + 4 tower layers
    assert line['order_first'][0].startswith(('cls_out', 'ins_out'))     # the loss-side parameters open the flat buffer
    for vals in line['vals']:
        for i, (n, v) in enumerate(zip(line['names'], vals)):
            want = 1.5 * (i + 1) if not n.startswith('ins_out') else 0.5 * (i + 1)    # mean over ranks; unused on rank 1 -> (g + 0) / 2
            assert abs(v - want) < 1e-6, (n, v, want)
    assert line['nbytes'] > 0


LOG_WORKER = r'''
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pointtinybenchmark_b200.dist import parse_losses
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
w = torch.ones(3, requires_grad=True)
losses = dict(gt_loss=(w * (1.0 + rank)).sum(), pos_loss=[w[0] * 2.0, w[1] * (4.0 + 2 * rank)], bag_acc=torch.tensor(0.5 + 0.25 * rank))
loss, log = parse_losses(losses)
loss.backward()
if rank == 0:
    print(json.dumps(dict(loss=float(loss), log=log, grad=w.grad.tolist(), keys=list(log))))
dist.destroy_process_group()
'''


def test_parse_losses_packed_allreduce_two_ranks(tmp_path):
    """logging reduction of detectors/base.py:179-212 as one packed all-reduce: the backward loss stays rank-local, the logged values
    are the mean over the ranks, keys without 'loss' (bag_acc) are logged but not summed"""
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    w = tmp_path / 'worker.py'
    w.write_text(LOG_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(w), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = json.loads([l for l in outs[0][0].splitlines() if l.startswith('{')][-1])
    assert line['keys'] == ['gt_loss', 'pos_loss', 'bag_acc', 'loss']
    assert line['loss'] == 3.0 + 6.0                                   # rank 0's own loss: gt 3 + pos (2 + 4)
    assert line['grad'] == [3.0, 5.0, 1.0]
    assert line['log'] == dict(gt_loss=(3.0 + 6.0) / 2, pos_loss=(6.0 + 8.0) / 2, bag_acc=0.625, loss=(9.0 + 14.0) / 2)


def test_reference_arm_contract():
    """--impl reference prints one JSON line with impl=reference and the e2e/cpu_baseline keys (CPU only, 1 step)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'img/s' and line['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['cpu_baseline']['kind'] == 'port'
    # non-zero ranks of a torchrun launch exit silently
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1'],
                       env=dict(os.environ, RANK='1', WORLD_SIZE='2'), capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ''
