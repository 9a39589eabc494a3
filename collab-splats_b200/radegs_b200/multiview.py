"""Camera-sharded multi-view training step.

The only data-parallel axis of the path is "independent cameras": every stage is per-camera and only the
per-Gaussian *gradients* couple views.  Gaussians are replicated on every GPU, rank r renders views
{r, r+G, ...} of the step's batch, and ONE ``all_reduce(sum)`` over a single flat fp32 gradient buffer per step
makes the replicas agree (NCCL over NVLink on the B200 box, gloo in the CPU tests).  Densification statistics
(per-Gaussian |grad2d| sums, visibility counts, max radii) are per-camera as well and need the same treatment
or the replicas diverge when they densify (last hard part).

Nothing here touches kernels: it is host-side plumbing over ``torch.distributed``.
"""

from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Views rendered by `rank`: round-robin so every rank gets ceil or floor(n_views / world) of them."""
    return list(range(rank, n_views, world))


class FlatGradBucket:
    """One contiguous fp32 buffer holding the gradients of all Gaussian parameters, in a fixed order.

    ``pack()`` copies ``p.grad`` of every parameter into the buffer (missing gradients count as zero),
    ``all_reduce()`` sums it across ranks, ``unpack()`` points every ``p.grad`` at its slice of the buffer
    (views, no copy).  Payload for 3 M Gaussians: 168 MB at sh0 (14 floats), 708 MB at sh3 (59 floats).
    """

    def __init__(self, params: Dict[str, Tensor]):
        self.names = list(params.keys())
        self.shapes = [tuple(params[k].shape) for k in self.names]
        self.sizes = [int(params[k].numel()) for k in self.names]
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + s)
        first = params[self.names[0]]
        self.buffer = torch.zeros(self.offsets[-1], dtype=torch.float32, device=first.device)

    def slices(self) -> Dict[str, Tensor]:
        return {k: self.buffer[o:o + n].view(shape)
                for k, o, n, shape in zip(self.names, self.offsets, self.sizes, self.shapes)}

    @torch.no_grad()
    def pack(self, params: Dict[str, Tensor]):
        views = self.slices()
        for k in self.names:
            g = params[k].grad
            if g is None:
                views[k].zero_()
            elif g.data_ptr() != views[k].data_ptr():
                views[k].copy_(g)
        return self.buffer

    def all_reduce(self, group=None, async_op: bool = False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None

    @torch.no_grad()
    def unpack(self, params: Dict[str, Tensor]):
        views = self.slices()
        for k in self.names:
            params[k].grad = views[k]


@torch.no_grad()
def sync_strategy_state(state: Dict, group=None):
    """Make the densification statistics identical on all ranks: sums for grad2d / count, max for radii."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for key, op in (("grad2d", dist.ReduceOp.SUM), ("count", dist.ReduceOp.SUM), ("radii", dist.ReduceOp.MAX)):
        t = state.get(key)
        if isinstance(t, Tensor):
            dist.all_reduce(t, op=op, group=group)


def global_batch_loss_scale(n_views: int) -> float:
    """Each rank averages the loss over the whole batch so that the summed gradient equals the gradient of the
    mean loss over all views of the step."""
    return 1.0 / float(n_views)


class ShGradExchange:
    """SH-coefficient gradients for the camera-sharded step WITHOUT all-reducing them (csrc/colors.cu, csrc/peer.cu).

    The coefficient gradient of one camera is the outer product ``Y_k(dir(n, camera)) x v_rgb[n]``: 48 floats per
    Gaussian that carry 3.  Each rank therefore publishes its clamp/visibility-masked colour gradients (12 B per
    Gaussian and camera) and every rank rebuilds the sum over the cameras of ALL ranks with one kernel
    (``rs_sh_coeffs_gather``), in a fixed (rank, camera) order, so all replicas hold bit-identical gradients.  At sh3
    this replaces a 192 B/Gaussian all-reduce (ring traffic 2(G-1)/G x 192 B) by (G-1) x 12 B/Gaussian of reads.

    ``mode="push"`` (default): every rank owns an inbox per rank in CUDA-IPC-shared device memory; right after the
    colour backward a rank copies its region into its slot of every peer's inbox with the COPY ENGINES (no SMs), on
    side streams, while the SMs run the projection VJP, then raises its flag at the peers; the gather kernel waits
    for the flags and reads only local memory.  ``mode="p2p"``: no copies, the gather kernel reads the other ranks'
    rows directly over NVLink (pull).  In both, a flag handshake (remote store / local spin, ``rs_peer_signal`` /
    ``rs_peer_wait``) orders publication before consumption, and two regions alternate so a rank may start its next
    backward while a slower peer still reads the previous one.  ``mode="allgather"``: the regions are ordinary tensors, all-gathered
    with ``torch.distributed`` (NCCL, or gloo-free single process), then the same kernel reads the local copy.

    Usage per step::

        ex.begin_step()
        with ex:                      # routes the backward of the fused SH colours to rs_sh_colors_bwd_local
            loss.backward()
        params["sh_coeffs"].grad = ex.finish()
    """

    def __init__(self, n_gaussians: int, cams_per_rank: int, device, group=None, mode: str = "push",
                 push_engine: str = "sm", push_ctas: int = 0,
                 timeout_ms: int = 5000):
        from radegs_b200 import backend as be
        import ctypes
        self.be, self.ct = be, ctypes
        self.lib = be.load()
        self.N, self.C = int(n_gaussians), int(cams_per_rank)
        self.device = torch.device(device)
        self.group = group
        self.timeout_ms = int(timeout_ms)
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        if self.world > 16:
            raise NotImplementedError("ShGradExchange supports up to 16 ranks (one NVSwitch domain)")
        self.region_bytes = int(self.lib.rs_sh_region_bytes(self.C, self.N))   # what THIS rank publishes per step
        self.step = 0
        self._pending = None
        self._out = None
        self.mode = mode
        if push_engine not in ("dma", "sm"):
            raise ValueError("push_engine must be 'dma' (copy engines) or 'sm' (store kernel)")
        # CTAs per peer of the SM store kernel; 0 = automatic.  The pushes overlap the projection VJP, so their CTAs
        # take SMs from it, while the bandwidth they need is bounded by the number of peers: about 32 pushing CTAs in
        # total is the measured optimum -- 8 GPUs at 16 MB per peer: 4 per peer 2.60 ms/step, 8: 2.68, 16: 2.70, copy
        # engines 2.74; 2 GPUs (one peer, which 4 CTAs cannot saturate): 4: 2.42 ms, 8-32: 2.39; at 48 MB per peer
        # (config 4) 8 per peer win at 8 GPUs (profiles/r02_exchange_ab_n8.txt, r02_push_ctas_n2.txt)
        self.push_engine = push_engine
        if int(push_ctas) > 0:
            self.push_ctas = int(push_ctas)
        else:
            peers = max(self.world - 1, 1)
            floor = 4 if self.region_bytes <= 24 * 2 ** 20 else 8
            self.push_ctas = max(floor, min(16, 32 // peers))
        self._peer_ptrs = []        # imported mappings, closed in close()
        self._own = []              # own cudaMalloc'ed blocks
        cams = [self.C] * self.world
        if ready and self.world > 1:
            got = [None] * self.world
            dist.all_gather_object(got, self.C, group=group)
            cams = [int(c) for c in got]
        self.cams = cams
        # Every rank derives the slot stride (and the inbox allocation) from the LARGEST per-rank camera count, so that
        # a sender's offset into a peer's inbox and the receiver's own view of it agree when the views of a step do not
        # divide evenly (shard_views gives ceil and floor shards); a rank still copies only its own region_bytes.
        self.region_stride = (int(self.lib.rs_sh_region_bytes(max(cams), self.N)) + 255) // 256 * 256
        with torch.cuda.device(self.device):
            if mode in ("p2p", "push"):
                self._init_p2p()
            elif mode == "allgather":
                if len(set(cams)) != 1:
                    raise NotImplementedError("allgather mode needs the same number of cameras on every rank")
                self.local = torch.zeros(2, self.region_stride, device=self.device, dtype=torch.uint8)
                self.gathered = torch.zeros(self.world, self.region_stride, device=self.device, dtype=torch.uint8)
            else:
                raise ValueError(mode)
            self.timed_out = torch.zeros(1, device=self.device, dtype=torch.int32)

    # ---- set-up: allocate, export, exchange handles, import
    def _alloc(self, nbytes: int) -> int:
        p = self.ct.c_void_p()
        self.be.check(self.lib.rs_peer_alloc(nbytes, self.ct.byref(p)), "rs_peer_alloc")
        self._own.append(p.value)
        return p.value

    def _init_p2p(self):
        ct, lib, be = self.ct, self.lib, self.be
        # slot (g, parity) at offset (2 g + parity) * stride: rank g's region of the steps with that parity.  In pull
        # mode only a rank's own slots are ever written; in push mode the peers fill the others by DMA.
        regions = self._alloc(2 * self.world * self.region_stride)
        flags = self._alloc(256)                               # u64[world], written by the peers
        hb = lib.rs_peer_handle_bytes()
        handles = []
        for ptr in (regions, flags):
            buf = ct.create_string_buffer(hb)
            be.check(lib.rs_peer_export(ct.c_void_p(ptr), buf), "rs_peer_export")
            handles.append(bytes(buf.raw))
        everyone = [handles]
        if self.world > 1:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handles, group=self.group)
        self.region_base, self.flag_base = [], []
        for g, (h_reg, h_flag) in enumerate(everyone):
            if g == self.rank:
                self.region_base.append(regions)
                self.flag_base.append(flags)
                continue
            mapped = []
            for h in (h_reg, h_flag):
                p = ct.c_void_p()
                be.check(lib.rs_peer_import(ct.create_string_buffer(h, hb), ct.byref(p)), "rs_peer_import")
                self._peer_ptrs.append(p.value)
                mapped.append(p.value)
            self.region_base.append(mapped[0])
            self.flag_base.append(mapped[1])
        self.flag_ptrs_dev = torch.tensor(self.flag_base, dtype=torch.int64, device=self.device)
        self._copy_streams = [torch.cuda.Stream(self.device) for _ in range(min(4, max(self.world - 1, 1)))]
        if self.world > 1:
            dist.barrier(group=self.group)                     # every mapping exists before anyone signals

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_available() and dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)                     # nobody still reads a region that is about to go
        for p in self._peer_ptrs:
            self.lib.rs_peer_unimport(self.ct.c_void_p(p))
        for p in self._own:
            self.lib.rs_peer_free(self.ct.c_void_p(p))
        self._peer_ptrs, self._own = [], []

    # ---- per step
    def begin_step(self):
        self.step += 1
        self._pending = None

    def __enter__(self):
        import gsplat.cuda._wrapper as wr
        self._prev_sink, wr.SH_GRAD_SINK = wr.SH_GRAD_SINK, self
        return self

    def __exit__(self, *exc):
        import gsplat.cuda._wrapper as wr
        wr.SH_GRAD_SINK = self._prev_sink

    def _parity(self) -> int:
        return self.step & 1

    def local_region_ptr(self, C: int, N: int):
        """Called by the SH-colour backward: where this rank's rows of the current step go."""
        if (C, N) != (self.C, self.N):
            raise RuntimeError(f"ShGradExchange was built for {self.C} cameras x {self.N} Gaussians, got {C} x {N}")
        if self._pending is not None:
            raise RuntimeError("one fused SH-colour backward per step (call begin_step() before the next one)")
        if self.mode in ("p2p", "push"):
            return self.ct.c_void_p(self.region_base[self.rank] + self._slot(self.rank))
        return self.be.ptr(self.local[self._parity()])

    def _slot(self, g: int) -> int:
        return (2 * g + self._parity()) * self.region_stride

    def published(self, means: Tensor, degree: int, K: int, stream_ptr):
        """Called right after rs_sh_colors_bwd_local was queued: tell the peers (p2p) and remember the inputs.
        (Queuing the wait + gather kernels here on a side stream, to overlap the projection VJP, was measured SLOWER
        at N=2 -- 2.66 vs 2.58 ms/step: that kernel fills the register file, so the two only take turns.)"""
        self._pending = (means, int(degree), int(K))
        be, lib, ct = self.be, self.lib, self.ct
        if self.mode == "p2p":
            be.check(lib.rs_peer_signal(be.ptr(self.flag_ptrs_dev), self.world, self.rank, self.step, stream_ptr),
                     "rs_peer_signal")
        elif self.mode == "push":
            main = torch.cuda.current_stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            src = self.region_base[self.rank] + self._slot(self.rank)
            peers = [g for g in range(self.world) if g != self.rank]
            streams = self._copy_streams
            for s_ in streams:
                s_.wait_event(fork)
            if self.push_engine == "sm" and peers:
                # one kernel, a few CTAs per peer, 16-byte posted stores over NVLink (rs_peer_push) on a side stream
                order = [peers[(i + self.rank) % len(peers)] for i in range(len(peers))]
                dsts = (ct.c_void_p * len(order))(*[self.region_base[g] + self._slot(self.rank) for g in order])
                with torch.cuda.stream(streams[0]):
                    be.check(lib.rs_peer_push(ct.cast(dsts, ct.c_void_p), len(order), ct.c_void_p(src),
                                              self.region_bytes, self.push_ctas, be.stream_ptr(self.device)),
                             "rs_peer_push")
            else:
                for i, g in enumerate(peers):      # start with the next rank so the ranks do not all hit rank 0 first
                    g = peers[(i + self.rank) % len(peers)]
                    with torch.cuda.stream(streams[i % len(streams)]):
                        be.check(lib.rs_peer_copy(ct.c_void_p(self.region_base[g] + self._slot(self.rank)),
                                                  ct.c_void_p(src), self.region_bytes, be.stream_ptr(self.device)),
                                 "rs_peer_copy")
                for s_ in streams[1:]:
                    ev = torch.cuda.Event()
                    ev.record(s_)
                    streams[0].wait_event(ev)
            with torch.cuda.stream(streams[0]):    # the flag goes up only after every copy has landed
                be.check(lib.rs_peer_signal(be.ptr(self.flag_ptrs_dev), self.world, self.rank, self.step,
                                            be.stream_ptr(self.device)), "rs_peer_signal")

    def _out_buffer(self, K: int) -> Tensor:
        """Two persistent result buffers alternate (the previous step's gradient stays valid for one more step)."""
        if self._out is None or self._out[0].shape[1] != K:
            self._out = [torch.empty(self.N, K, 3, device=self.device, dtype=torch.float32) for _ in range(2)]
        return self._out[self._parity()]

    @torch.no_grad()
    def finish(self) -> Tensor:
        """-> d loss / d sh_coeffs [N,K,3] summed over the cameras of all ranks (identical on every rank).  The
        returned tensor is one of two buffers owned by the exchange: it stays valid until the step after next."""
        if self._pending is None:
            raise RuntimeError("finish() without a backward through the fused SH colours in this step")
        means, degree, K = self._pending
        ct, lib, be = self.ct, self.lib, self.be
        out = self._out_buffer(K)
        with torch.cuda.device(self.device):
            st = be.stream_ptr(self.device)
            if self.mode in ("p2p", "push"):
                be.check(lib.rs_peer_wait(ct.c_void_p(self.flag_base[self.rank]), self.world, self.step,
                                          self.timeout_ms, be.ptr(self.timed_out), st), "rs_peer_wait")
                if self.mode == "p2p":      # pull: rank g's slot in rank g's memory
                    bases = [self.region_base[g] + self._slot(g) for g in range(self.world)]
                else:                       # push: rank g's slot in MY inbox
                    bases = [self.region_base[self.rank] + self._slot(g) for g in range(self.world)]
            else:
                if self.world > 1:
                    dist.all_gather_into_tensor(self.gathered.view(-1), self.local[self._parity()], group=self.group)
                else:
                    self.gathered[0].copy_(self.local[self._parity()])
                base = self.gathered.data_ptr()
                bases = [base + g * self.region_stride for g in range(self.world)]
            regions = (ct.c_void_p * self.world)(*bases)
            cams = (ct.c_int * self.world)(*self.cams)
            be.check(lib.rs_sh_coeffs_gather(degree, K, self.N, be.ptr(means), regions, cams, self.world, be.ptr(out),
                                             st), "rs_sh_coeffs_gather")
        self._pending = None
        return out

    def check(self):
        """Raises if a peer failed to publish within the timeout (one device->host read: call it off the hot path)."""
        if int(self.timed_out.item()) != 0:
            raise RuntimeError("ShGradExchange: a peer rank did not publish its colour gradients in time")


class _RawCudaBuffer:
    """Lets torch view a raw device pointer (peer-visible memory from rs_peer_alloc) as a tensor."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerAllReduce:
    """Sum of one flat fp32 vector over the ranks WITHOUT NCCL: the two-shot all-reduce kernel of csrc/peer.cu
    (``rs_peer_allreduce``) over CUDA-IPC-mapped buffers, bracketed by the flag handshake.

    ``self.flat`` is a torch view of this rank's peer-visible buffer: write the local values into it (e.g.
    ``torch.cat(grads, out=ar.flat[:n])``), call ``all_reduce()``, read the sums from the same tensor.  Every rank adds
    the contributions in rank order, once per slice, so all replicas hold bit-identical results.  Per step each rank
    moves (G-1)/G of the vector in over NVLink and the same amount out (the ring all-reduce's volume) in one kernel.
    """

    def __init__(self, n_floats: int, device, group=None, ctas: int = 296, timeout_ms: int = 5000):
        from radegs_b200 import backend as be
        import ctypes
        self.be, self.ct, self.lib = be, ctypes, be.load()
        self.device = torch.device(device)
        self.group = group
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.rank = dist.get_rank(group) if ready else 0
        if self.world not in (1, 2, 4, 8):
            raise NotImplementedError("PeerAllReduce supports 1, 2, 4 or 8 ranks")
        quantum = 4 * self.world
        self.n = (int(n_floats) + quantum - 1) // quantum * quantum
        self.ctas, self.timeout_ms = int(ctas), int(timeout_ms)
        self.step = 0
        self._peer_ptrs, self._own = [], []
        ct, lib = ctypes, self.lib
        with torch.cuda.device(self.device):
            ptrs = []
            for nbytes in (4 * self.n, 256, 256):            # data, "inputs ready" flags, "sums landed" flags
                p = ct.c_void_p()
                be.check(lib.rs_peer_alloc(nbytes, ct.byref(p)), "rs_peer_alloc")
                self._own.append(p.value)
                ptrs.append(p.value)
            hb = lib.rs_peer_handle_bytes()
            handles = []
            for ptr in ptrs:
                buf = ct.create_string_buffer(hb)
                be.check(lib.rs_peer_export(ct.c_void_p(ptr), buf), "rs_peer_export")
                handles.append(bytes(buf.raw))
            everyone = [handles]
            if self.world > 1:
                everyone = [None] * self.world
                dist.all_gather_object(everyone, handles, group=group)
            self.bases = [[], [], []]
            for g, hs in enumerate(everyone):
                for k, h in enumerate(hs):
                    if g == self.rank:
                        self.bases[k].append(ptrs[k])
                        continue
                    p = ct.c_void_p()
                    be.check(lib.rs_peer_import(ct.create_string_buffer(h, hb), ct.byref(p)), "rs_peer_import")
                    self._peer_ptrs.append(p.value)
                    self.bases[k].append(p.value)
            self.flat = torch.as_tensor(_RawCudaBuffer(ptrs[0], self.n), device=self.device)
            self.ready_dev = torch.tensor(self.bases[1], dtype=torch.int64, device=self.device)
            self.done_dev = torch.tensor(self.bases[2], dtype=torch.int64, device=self.device)
            self.timed_out = torch.zeros(1, device=self.device, dtype=torch.int32)
        if self.world > 1:
            dist.barrier(group=group)

    @torch.no_grad()
    def all_reduce(self) -> Tensor:
        """Sums ``self.flat`` over the ranks in place, on the current stream; returns ``self.flat``."""
        if self.world == 1:
            return self.flat
        be, lib, ct = self.be, self.lib, self.ct
        self.step += 1
        with torch.cuda.device(self.device):
            st = be.stream_ptr(self.device)
            be.check(lib.rs_peer_signal(be.ptr(self.ready_dev), self.world, self.rank, self.step, st), "rs_peer_signal")
            be.check(lib.rs_peer_wait(ct.c_void_p(self.bases[1][self.rank]), self.world, self.step, self.timeout_ms,
                                      be.ptr(self.timed_out), st), "rs_peer_wait")
            bufs = (ct.c_void_p * self.world)(*self.bases[0])
            be.check(lib.rs_peer_allreduce(bufs, self.world, self.rank, self.n, self.ctas, st), "rs_peer_allreduce")
            be.check(lib.rs_peer_signal(be.ptr(self.done_dev), self.world, self.rank, self.step, st), "rs_peer_signal")
            be.check(lib.rs_peer_wait(ct.c_void_p(self.bases[2][self.rank]), self.world, self.step, self.timeout_ms,
                                      be.ptr(self.timed_out), st), "rs_peer_wait")
        return self.flat

    def check(self):
        if int(self.timed_out.item()) != 0:
            raise RuntimeError("PeerAllReduce: a peer rank did not arrive in time")

    def close(self):
        torch.cuda.synchronize(self.device)
        if dist.is_available() and dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)
        self.flat = None
        for p in self._peer_ptrs:
            self.lib.rs_peer_unimport(self.ct.c_void_p(p))
        for p in self._own:
            self.lib.rs_peer_free(self.ct.c_void_p(p))
        self._peer_ptrs, self._own = [], []
