"""Row-sharded match on real GPUs.  world=1 runs everywhere; the 2-rank NCCL case needs >= 2 devices
(it is launched as its own torchrun job) and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import matcher_oracle as mo
from oracle import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_world1_equals_plain_matcher():
    import torch
    import facerecognition_infrenceengine_b200 as frg
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
    n, d, f, k = 50000, 512, 24, 5
    store = frg.GalleryStore(dim=d, capacity=n)
    g = ShardedGallery(dim=d, device=0, store=store, rank=0, world=1)
    g.fill_synthetic(n, 1234)
    Q, _ = synth.queries(f, n, d)
    rows, scores, acc = ShardedMatcher(g).match(torch.from_numpy(Q).cuda(), k, 0.45)
    torch.cuda.synchronize()
    ref = frg.Matcher(store).match(Q, k, 0.45)
    assert (rows.cpu().numpy() == ref.rows).all() and (scores.cpu().numpy() == ref.scores).all()
    assert (acc.cpu().numpy().astype(bool) == ref.accept).all()
    store.close()


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import facerecognition_infrenceengine_b200 as frg
from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
from oracle import matcher_oracle as mo, synth
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, f, k = 300001, 512, 64, 10
store = frg.GalleryStore(dim=d, capacity=n, device=local)
g = ShardedGallery(dim=d, device=local, store=store)
g.fill_synthetic(n, 1234)
Q, target = synth.queries(f, n, d)
Qd = torch.from_numpy(Q).cuda()
m_nccl = ShardedMatcher(g, exchange="nccl")
rows, scores, acc = m_nccl.match(Qd, k, 0.45)
torch.cuda.synchronize()
assert m_nccl.exchange == "nccl"
if dist.get_rank() == 0:
    G = synth.gallery(n, d)
    rr, rs, ra = mo.match_topk(Q, G, k + 1, 0.45)
    assert mo.ids_match_with_gap(rr, rs, rows.cpu().numpy(), 1e-4).all()
    assert np.abs(scores.cpu().numpy() - rs[:, :k]).max() <= 1e-4
    assert (acc.cpu().numpy().astype(bool) == ra).all()
    print("SHARDED_OK world=%%d" %% dist.get_world_size())
# the fused peer-memory exchange (push + flag + merge in one kernel, no collective call) gives the same
# answer bit for bit, call after call (epoch parity reuse), for changing batch sizes (buffer regrowth), and
# when one rank runs late
m_p2p = ShardedMatcher(g, exchange="p2p")
for it, (ff, kk) in enumerate([(64, 10), (64, 10), (64, 10), (8, 1), (200, 5), (64, 10), (2, 16)] * 3):
    Qi = Qd[:ff] if ff <= f else torch.from_numpy(synth.queries(ff, n, d, seed=77 + it)[0]).cuda()
    if it %% 4 == dist.get_rank():
        torch.cuda._sleep(200_000_000)              # ~0.1 s of GPU time: this rank arrives late
    a = m_p2p.match(Qi, kk, 0.45)
    b = m_nccl.match(Qi, kk, 0.45)
    torch.cuda.synchronize()
    assert m_p2p.exchange == "p2p", m_p2p.p2p_error
    for x, y in zip(a, b):
        assert torch.equal(x, y), (it, ff, kk)
# variants without a select stage: the exchange kernel pushes every query itself
a = m_p2p.match(Qd, 5, 0.45, variant="scan_f32"); b = m_nccl.match(Qd, 5, 0.45, variant="scan_f32")
torch.cuda.synchronize()
for x, y in zip(a, b):
    assert torch.equal(x, y)
# adversarial shard: 3000 copies of one template land on the LAST rank; the queries near it overflow that
# rank's candidate lists, are redone by its exact fallback, and reach the other ranks through the late push
dupe = synth.unit_rows(np.arange(1), d, 4242, synth.STREAM_IMPOSTOR)
g.append_local(np.repeat(dupe, 3000, axis=0), prenormalised=True)
Qx = torch.from_numpy(np.concatenate([dupe, dupe + np.float32(1e-3), Q[:14]])).cuda()
for kk in (1, 16):
    a = m_p2p.match(Qx, kk, 0.45); b = m_nccl.match(Qx, kk, 0.45)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y), kk
    assert a[0][0].tolist() == list(range(n, n + kk)), a[0][0].tolist()      # ties -> earliest rows, global numbering
# every rank holds the same merged result
chk = a[0].clone()
dist.broadcast(chk, src=0)
assert torch.equal(chk, a[0])
if dist.get_rank() == 0:
    print("P2P_OK world=%%d epochs=%%d" %% (dist.get_world_size(), m_p2p._epoch))
dist.destroy_process_group()
'''


def test_sharded_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK world=2" in r.stdout
    assert "P2P_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
