"""Expert-parallel exchange for the routed MoE experts: one process per GPU, NCCL over NVLink.

The reference keeps all 64 experts on one device (modeling_bailing_moe.py:543-549); there is nothing to mirror.  Here the
experts are sharded (rank r owns experts [r E/G, (r+1) E/G)) and, for prefill-sized inputs, so are the tokens of the MoE
block: every rank routes its slice of the tokens, the (token, slot) rows travel to the rank that owns their expert in ONE
all-to-all, come back in a second all-to-all after the grouped expert GEMMs, are combined in fp32 at the owner (the
reference's summation order, :632-638), and the slices are all-gathered for the replicated attention of the next layer.

`torch.distributed` is the plumbing (all_to_all_single / all_gather_into_tensor); the row layouts on both sides are
produced by the CUDA routing-plan kernel (mb_moe_plan with granule 1: rows ordered by destination rank).  The class is
device- and dtype-agnostic so the world-size-2 gloo test can drive it on CPU tensors.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ExpertParallelAllToAll:
    """All-to-all dispatch / return of variable-length row blocks between the ranks of `group`."""

    def __init__(self, group=None):
        self.group = group
        self.size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def exchange_counts(self, send_counts: torch.Tensor) -> tuple[list[int], list[int]]:
        """send_counts int32 [G] (rows for every destination) -> (send_splits, recv_splits) as host lists.
        One tiny all-to-all plus ONE host read per call (the reference reads the per-expert counts on the host in
        every layer too, modeling_bailing_moe.py:616)."""
        if send_counts.numel() != self.size:
            raise ValueError(f"send_counts must have {self.size} entries")
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        both = torch.stack([send_counts, recv_counts]).tolist()
        return [int(v) for v in both[0]], [int(v) for v in both[1]]

    def dispatch(self, send_rows: torch.Tensor, send_ids: torch.Tensor, send_splits: list[int],
                 recv_splits: list[int]) -> tuple[torch.Tensor, torch.Tensor]:
        """send_rows [n, D] ordered by destination rank (send_splits rows each) with their expert ids send_ids [n] ->
        (recv_rows [m, D], recv_ids [m]) ordered by source rank."""
        if send_rows.shape[0] != sum(send_splits) or send_ids.shape[0] != send_rows.shape[0]:
            raise ValueError("send buffer does not match the split sizes")
        m = sum(recv_splits)
        recv_rows = send_rows.new_empty((m, send_rows.shape[1]))
        recv_ids = send_ids.new_empty((m,))
        dist.all_to_all_single(recv_rows, send_rows.contiguous(), recv_splits, send_splits, group=self.group)
        dist.all_to_all_single(recv_ids, send_ids.contiguous(), recv_splits, send_splits, group=self.group)
        return recv_rows, recv_ids

    def combine(self, out_rows: torch.Tensor, send_splits: list[int], recv_splits: list[int]) -> torch.Tensor:
        """The way back: out_rows [m, D] in arrival order of `dispatch` -> [n, D] in the order the rows were sent."""
        if out_rows.shape[0] != sum(recv_splits):
            raise ValueError("returned rows do not match the split sizes")
        back = out_rows.new_empty((sum(send_splits), out_rows.shape[1]))
        dist.all_to_all_single(back, out_rows.contiguous(), send_splits, recv_splits, group=self.group)
        return back

    def all_gather_rows(self, local_rows: torch.Tensor) -> torch.Tensor:
        """local_rows [Tc, D] (same Tc on every rank) -> [G * Tc, D] in rank order."""
        out = local_rows.new_empty((self.size * local_rows.shape[0], local_rows.shape[1]))
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=self.group)
        return out


def token_slice(T: int, size: int, rank: int) -> tuple[int, int, int]:
    """Contiguous token slice of `rank`: chunk = ceil(T / size); returns (begin, end, chunk) with end <= T."""
    chunk = (T + size - 1) // size
    return min(rank * chunk, T), min((rank + 1) * chunk, T), chunk


class PeerExchange:
    """Exchange area of the FUSED expert-parallel combine (mb_moe_combine_push / mb_moe_reduce_finalize): one buffer
    per rank in torch symmetric memory (cuMem allocation mapped into every process of the group over NVLink), plus the
    device array of the peers' base pointers that the kernels store through.  One instance serves every MoE layer of a
    model: the calls are stream-ordered and the kernels' epoch / parity protocol keeps consecutive calls apart."""

    T_MAX = 8  # rows per call (decode regime: CFG rows)

    def __init__(self, group, hidden_size: int, device):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        self.group = group if group is not None else dist.group.WORLD
        self.size, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.hidden_size = hidden_size
        import ctypes

        nbytes = ctypes.c_int64(0)
        _lib.check(_lib.load().mb_moe_peer_area_bytes(self.size, self.T_MAX, hidden_size, ctypes.byref(nbytes)),
                   "mb_moe_peer_area_bytes")
        n = (nbytes.value + 3) // 4
        self.buf = symm_mem.empty((n,), dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)  # collective: exchanges the memory handles
        if len(self.handle.buffer_ptrs) != self.size:
            raise RuntimeError("symmetric-memory rendezvous did not return one buffer per rank")
        self.peers_dev = int(self.handle.buffer_ptrs_dev)  # device array of G base pointers (as mapped in this process)
        self.fin_done = torch.zeros((1,), dtype=torch.int32, device=device)
        dist.barrier(self.group)  # every area is zeroed before anyone pushes


class PeerDispatch:
    """Exchange areas of the peer-memory DISPATCH + COMBINE (csrc/ep.cu: mb_ep_dispatch_push / _wait / mb_ep_combine_push /
    mb_ep_reduce_finalize) — "data parallel x expert parallel": every rank runs its OWN token rows (its own image or
    request) through replicated attention / gates / shared experts and owns E / G routed experts; per MoE layer the rows
    are all-gathered through NVLink peer stores, each rank runs its experts on the rows of all ranks, and the fp32
    partial sums travel back to their owners the same way.  No NCCL call and no host synchronisation per layer, so a
    whole token step (28 layers) stays ONE CUDA graph.

    One instance serves every MoE layer of a model (calls are stream-ordered; the device-side epoch keeps them apart).
    `t_max` = most rows per rank and call (CFG rows at decode, prompt length at prefill); every rank must make the same
    calls with the same T.  Built either over torch symmetric memory (one process per GPU, `group`) or — `local_ranks=G`
    — as G virtual ranks on ONE device, which is how the protocol is tested on a single-GPU box."""

    def __init__(self, group, hidden_size: int, top_k: int, t_max: int, device, local_ranks: int = 0):
        import ctypes

        from . import _lib

        lib = _lib.load()
        self.hidden_size, self.top_k, self.t_max = hidden_size, top_k, t_max
        self.local = local_ranks > 0
        if self.local:
            self.group, self.size, self.rank = None, local_ranks, 0
        else:
            self.group = group if group is not None else dist.group.WORLD
            self.size, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        offs = (ctypes.c_int64 * 6)()
        _lib.check(lib.mb_ep_area_layout(self.size, t_max, hidden_size, top_k, offs), "mb_ep_area_layout")
        self.off_x, self.off_idx, self.off_w, self.off_part, self.off_ctrl, self.nbytes = (int(v) for v in offs)
        if self.local:
            self.bufs = [torch.zeros((self.nbytes,), dtype=torch.uint8, device=device) for _ in range(self.size)]
            self._ptrs = torch.tensor([b.data_ptr() for b in self.bufs], dtype=torch.int64, device=device)
            self.peers_dev = self._ptrs.data_ptr()
        else:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty((self.nbytes,), dtype=torch.uint8, device=device)
            buf.zero_()
            torch.cuda.synchronize(device)
            self.handle = symm_mem.rendezvous(buf, self.group)  # collective: exchanges the memory handles
            if len(self.handle.buffer_ptrs) != self.size:
                raise RuntimeError("symmetric-memory rendezvous did not return one buffer per rank")
            self.bufs = [None] * self.size
            self.bufs[self.rank] = buf
            self.peers_dev = int(self.handle.buffer_ptrs_dev)
            dist.barrier(self.group)  # every area is zeroed before anyone pushes

    def area(self, rank: int | None = None) -> torch.Tensor:
        return self.bufs[self.rank if rank is None else rank]

    def gathered(self, T: int, rank: int | None = None):
        """Views of the local area after mb_ep_dispatch_wait: (x [G*T, D] bf16, idx [G*T, k] i32, w [G*T, k] f32)."""
        a, G, D, k = self.area(rank), self.size, self.hidden_size, self.top_k
        x = a[self.off_x:self.off_x + G * T * D * 2].view(torch.bfloat16).view(G * T, D)
        idx = a[self.off_idx:self.off_idx + G * T * k * 4].view(torch.int32).view(G * T, k)
        w = a[self.off_w:self.off_w + G * T * k * 4].view(torch.float32).view(G * T, k)
        return x, idx, w

    def check(self, rank: int | None = None) -> None:
        """Raises if a bounded wait of this rank's kernels expired (call after a synchronisation point)."""
        G = self.size
        ctl = self.area(rank)[self.off_ctrl:self.off_ctrl + (2 * G + 8) * 4].view(torch.int32)
        code, peer = int(ctl[2 * G + 4]), int(ctl[2 * G + 5])
        if code != 0:
            raise RuntimeError(f"expert-parallel exchange timed out waiting for rank {peer} "
                               f"({'dispatch' if code == 1 else 'combine'} flag); the step's output is invalid")
