"""Expert-parallel dispatch + combine over peer memory (csrc/ep.cu; "data parallel x expert parallel") on ONE GPU: G
virtual ranks, each with its own exchange area, its own token rows and its own slab of E / G experts, driven phase by
phase on one stream exactly as G processes would run them (ep.PeerDispatch(local_ranks=G)).  Checks every rank's
output against the unsharded MoE block on the same rows, over several consecutive calls of varying row counts (epoch
protocol, single-buffered areas), the grouped-GEMM and the streaming expert paths, and a CUDA-graph replay of the whole
exchange.  The two-process version over NVLink runs in tests/test_ep_gpu.py (needs >= 2 GPUs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _setup(dev, G, D, E, k, I, seed=5):
    g = torch.Generator().manual_seed(seed)
    Wgu = (torch.randn((E, 2 * I, D), generator=g) / D ** 0.5).to(dev).to(BF16)
    Wd = (torch.randn((E, D, I), generator=g) / I ** 0.5).to(dev).to(BF16)
    n_local = E // G
    slabs = [(Wgu[r * n_local:(r + 1) * n_local].contiguous(), Wd[r * n_local:(r + 1) * n_local].contiguous())
             for r in range(G)]
    return g, Wgu, Wd, slabs, n_local


def _inputs(g, dev, T, D, E, k):
    x = torch.randn((T, D), generator=g).to(dev).to(BF16)
    res = torch.randn((T, D), generator=g).to(dev).to(BF16)
    sh = torch.randn((T, D), generator=g).to(dev).to(BF16)
    idx = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)]).to(torch.int32).to(dev)
    w = torch.rand((T, k), generator=g).to(dev)
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    return x, idx, w, sh, res


def _exchange(pd, ops, ins, slabs, n_local, E, T):
    G = pd.size
    for r in range(G):
        ops.ep_dispatch(pd, ins[r][0], ins[r][1], ins[r][2], rank=r)
    outs = [ops.ep_compute(pd, T, slabs[r][0], slabs[r][1], r * n_local, E, rank=r) for r in range(G)]
    for r in range(G):
        ops.ep_combine(pd, T, outs[r][0], outs[r][1], r * n_local, n_local, rank=r)
    return [ops.ep_finalize(pd, T, ins[r][3], ins[r][4], rank=r) for r in range(G)]


@pytest.mark.parametrize("G,D,E,k,I,Ts", [(2, 256, 16, 2, 64, (1, 3, 2, 8, 3, 40)),
                                          (8, 2048, 64, 6, 1408, (3, 3, 2, 64)),
                                          (4, 2048, 64, 6, 1408, (200,))])
def test_dispatch_combine_virtual_ranks(cuda_device, G, D, E, k, I, Ts):
    from ming_univision_b200 import ops
    from ming_univision_b200.ep import PeerDispatch

    dev = cuda_device
    g, Wgu, Wd, slabs, n_local = _setup(dev, G, D, E, k, I)
    pd = PeerDispatch(None, D, k, max(Ts), dev, local_ranks=G)
    for T in Ts:
        ins = [_inputs(g, dev, T, D, E, k) for _ in range(G)]
        ys = _exchange(pd, ops, ins, slabs, n_local, E, T)
        torch.cuda.synchronize()
        for r in range(G):
            pd.check(r)
            x, idx, w, sh, res = ins[r]
            ref = ops.moe_experts(x, idx, w, Wgu, Wd, sh, res)
            err = float((ys[r].float() - ref.float()).norm() / ref.float().norm())
            # identical expert outputs; only the ORDER of the fp32 weighted sum differs (per rank, then over ranks)
            assert err < 4e-3, (G, T, r, err)
            assert float((ys[r].float() - ref.float()).abs().max()) <= 2.0 ** -6 * float(ref.float().abs().max())
        # deterministic: the same call again gives the same bits
        ys2 = _exchange(pd, ops, ins, slabs, n_local, E, T)
        assert all(torch.equal(a, b) for a, b in zip(ys, ys2))


def test_dispatch_combine_in_a_cuda_graph(cuda_device):
    """No NCCL call, no host read, device-side epoch: the whole exchange of all ranks is capturable, and replays."""
    from ming_univision_b200 import ops
    from ming_univision_b200.ep import PeerDispatch

    dev = cuda_device
    G, D, E, k, I, T = 4, 2048, 64, 6, 1408, 3
    g, Wgu, Wd, slabs, n_local = _setup(dev, G, D, E, k, I)
    pd = PeerDispatch(None, D, k, 8, dev, local_ranks=G)
    ins = [_inputs(g, dev, T, D, E, k) for _ in range(G)]
    eager = _exchange(pd, ops, ins, slabs, n_local, E, T)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _exchange(pd, ops, ins, slabs, n_local, E, T)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ys = _exchange(pd, ops, ins, slabs, n_local, E, T)
    for it in range(3):
        if it == 2:  # new inputs in the static buffers
            new = [_inputs(g, dev, T, D, E, k) for _ in range(G)]
            for r in range(G):
                for a, b in zip(ins[r], new[r]):
                    a.copy_(b)
            eager = None
        graph.replay()
        torch.cuda.synchronize()
        for r in range(G):
            pd.check(r)
        if eager is not None:
            assert all(torch.equal(a, b) for a, b in zip(ys, eager))
    for r in range(G):
        x, idx, w, sh, res = ins[r]
        ref = ops.moe_experts(x, idx, w, Wgu, Wd, sh, res)
        assert float((ys[r].float() - ref.float()).norm() / ref.float().norm()) < 4e-3


def test_dispatch_wait_times_out_instead_of_hanging(cuda_device, monkeypatch):
    """A rank whose peer never arrives must get an error code back, not a hung (or trapped) GPU."""
    import subprocess
    import sys
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys; sys.path.insert(0, {root!r})
import torch
from ming_univision_b200 import ops
from ming_univision_b200.ep import PeerDispatch
dev = torch.device("cuda:0")
pd = PeerDispatch(None, 256, 2, 8, dev, local_ranks=2)
x = torch.zeros((3, 256), dtype=torch.bfloat16, device=dev)
idx = torch.zeros((3, 2), dtype=torch.int32, device=dev); w = torch.ones((3, 2), device=dev)
ops.ep_dispatch(pd, x, idx, w, rank=0)          # rank 1 never dispatches
from ming_univision_b200 import _lib
_lib.check(_lib.load().mb_ep_dispatch_wait(pd.peers_dev, 0, 2, 8, 256, 2, torch.cuda.current_stream().cuda_stream), "wait")
torch.cuda.synchronize()
try:
    pd.check(0)
    print("NO ERROR")
except RuntimeError as e:
    print("TIMEOUT REPORTED:", e)
y = torch.ones(4, device=dev) * 2; torch.cuda.synchronize(); print("context alive", float(y.sum()))
"""
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MB_EP_TIMEOUT_MS="200"), capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "TIMEOUT REPORTED" in r.stdout and "rank 1" in r.stdout and "context alive 8.0" in r.stdout, r.stdout
