"""2-GPU (NCCL box): the one-shot peer-memory all-reduce kernel against ``dist.all_reduce``, eager and inside a
replayed CUDA graph, P2P and (where the fabric offers it) NVSwitch multicast reduction.  Skipped on < 2 GPUs; the
N > 1 host logic is covered on CPU by tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DGGB_ROOT"])
import dgg_b200
from dgg_b200 import sharding as S
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
assert S.PeerAllReduce.available()
n = 36165
for use_mc in (False, True):
    ar = S.PeerAllReduce(n, use_multicast=use_mc)
    if use_mc and not ar.multicast:
        print("rank", rank, "no multicast object on this fabric: skipped", flush=True)
        continue
    for step in range(5):
        g = torch.Generator(device="cuda").manual_seed(100 * step + rank)
        ar.buffer.copy_(torch.randn(n, generator=g, device="cuda"))
        want = ar.buffer.clone()
        dist.all_reduce(want)
        got = ar().clone()
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        assert all(torch.equal(gathered[0], t) for t in gathered)          # identical bits on every rank
    # inside a CUDA graph: fill + all-reduce captured, replayed 20 times
    src = torch.randn(n, device="cuda")
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ar.buffer.copy_(src); ar()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        ar.buffer.copy_(src * 1.5)
        ar()
    want = src * 1.5
    dist.all_reduce(want)
    for _ in range(20):
        gr.replay()
    torch.testing.assert_close(ar.out, want, rtol=1e-6, atol=1e-6)
    print("rank", rank, "multicast" if ar.multicast else "p2p", "ok", flush=True)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_matches_nccl(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DGGB_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") >= 2
