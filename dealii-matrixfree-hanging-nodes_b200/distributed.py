"""Ghost exchange of the partitioned operator: the device side of
LinearAlgebra::distributed::Vector::update_ghost_values / compress(add) as
CUDAWrappers::MatrixFree::cell_loop uses them (reference: benchmark_03.h:348-353
with the distributed vectors of :323-324; deal.II interleaves the two exchanges
with three cell partitions).

One process per GPU; torch.distributed (NCCL over NVLink, gloo in the CPU
tests) carries the messages, this package's own kernels pack / unpack on the
GPU.  Schedule of one vmult (compute stream || communication stream):

    pack imports            |
    interior cells, part A  |  owners -> ghosts   (update_ghost_values)
    boundary cells          |
    interior cells, part B  |  ghosts -> owners   (compress, part 1)
    unpack-add, zero ghosts |
"""
from __future__ import annotations

import numpy as np

from . import _capi as capi
from ._capi import check, lib


class DeviceVector:
    """Device vector allocated by the library (cudaMalloc) so that it can be exported to the other
    ranks with CUDA IPC; exposes __cuda_array_interface__, `tensor` is a torch view of it."""

    def __init__(self, n, dtype, device, padding=0):
        import ctypes as C

        import torch

        self.n, self.itemsize, self.padding = int(n), torch.empty(0, dtype=dtype).element_size(), int(padding)
        self.typestr = "<f8" if self.itemsize == 8 else "<f4"
        p = C.c_void_p()
        check(lib.mfhn_vec_alloc((self.n + self.padding) * self.itemsize, C.byref(p)))  # zeroed
        self.ptr = p.value
        self.tensor = torch.as_tensor(self, device=device)[:self.n]  # the spare entries stay behind the view

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n + self.padding,), "typestr": self.typestr, "data": (self.ptr, False), "version": 2, "strides": None}

    def ipc_handle(self):
        import torch

        h = torch.zeros(64, dtype=torch.uint8)
        check(lib.mfhn_ipc_get_handle(self.ptr, h.data_ptr()))
        return h

    def __del__(self):
        if getattr(self, "ptr", None):
            self.tensor = None
            lib.mfhn_vec_free(self.ptr)
            self.ptr = None


class GhostExchange:
    """Owns the send/receive buffers and index lists of one rank.

    local_apply(dst, src, cell_begin, cell_end) applies the rank-local cell
    loop on a cell range; by default the CUDA operator's vmult_range.  The CPU
    tests pass the oracle here to exercise the exchange logic under gloo."""

    def __init__(self, op=None, partitioner=None, segments=None, local_apply=None, device=None, dtype=None, group=None,
                 native=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.op = op
        part = partitioner if partitioner is not None else op.mf.partitioner
        self.part = part
        self.device = device if device is not None else op.device
        self.dtype = dtype if dtype is not None else op.dtype
        self.number = capi.F64 if self.dtype == torch.float64 else capi.F32
        self.cuda = torch.device(self.device).type == "cuda"
        self.local_apply = local_apply if local_apply is not None else op.vmult_range
        # cell segments [interior A | interior B | boundary]
        if segments is None:
            mf = op.mf
            segments = (0, mf.n_interior_a, mf.n_interior_cells, mf.n_cells)
        self.seg = tuple(int(s) for s in segments)
        self.n_owned, self.n_ghost = part.n_owned, part.n_ghost
        self.import_peers = sorted(part.import_indices)
        self.ghost_peers = sorted(part.ghost_ranges)
        self.import_idx = {p: torch.from_numpy(np.ascontiguousarray(part.import_indices[p], dtype=np.int32)).to(self.device)
                           for p in self.import_peers}
        self.send_buf = {p: torch.empty(len(part.import_indices[p]), dtype=self.dtype, device=self.device) for p in self.import_peers}
        self.recv_buf = {p: torch.empty(len(part.import_indices[p]), dtype=self.dtype, device=self.device) for p in self.import_peers}
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        self.n_launches = 0
        # native path: the whole schedule (pack, NCCL groups, cell partitions, unpack) in one C-ABI call
        self._native = None
        if native is None:
            native = self.cuda and op is not None and local_apply is None and dist.get_backend(group) == "nccl"
        if native:
            self._create_native(op)

    def _create_native(self, op):
        import ctypes as C

        torch, dist, part = self.torch, self.dist, self.part
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            check(lib.mfhn_dist_unique_id(uid.data_ptr()))
        uid = uid.to(self.device)
        dist.broadcast(uid, src=self._global_rank(0), group=self.group)
        uid = uid.cpu().contiguous()
        h = C.c_void_p()
        mf = getattr(op, "mf", None)
        if getattr(mf, "_h", None) is not None and part is mf.partitioner:
            # the MatrixFree handle holds the partitioner and the cell partitions: one call
            check(lib.mfhn_dist_create_mf(op._h, mf._h, uid.data_ptr(), C.byref(h)))
        else:
            ip = np.array(self.import_peers, dtype=np.int32)
            ioff = np.zeros(len(ip) + 1, dtype=np.int64)
            for i, p in enumerate(self.import_peers):
                ioff[i + 1] = ioff[i] + len(part.import_indices[p])
            iidx = (np.concatenate([part.import_indices[p] for p in self.import_peers]).astype(np.int32) if len(ip)
                    else np.zeros(0, dtype=np.int32))
            gp = np.array(self.ghost_peers, dtype=np.int32)
            gb = np.array([part.ghost_ranges[p][0] for p in self.ghost_peers], dtype=np.int64)
            ge = np.array([part.ghost_ranges[p][1] for p in self.ghost_peers], dtype=np.int64)
            ptr = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
            desc = capi.DistDesc(rank=rank, world=world, unique_id=uid.data_ptr(), n_import_peers=len(ip), import_peers=ptr(ip),
                                 import_offsets=ptr(ioff), import_indices=ptr(iidx), n_ghost_peers=len(gp), ghost_peers=ptr(gp),
                                 ghost_begin=ptr(gb), ghost_end=ptr(ge), segments=(C.c_int64 * 4)(*self.seg))
            check(lib.mfhn_dist_create(op._h, C.byref(desc), C.byref(h)))
        self._native = h

    def __del__(self):
        for p in getattr(self, "_peer_opened", []):
            lib.mfhn_ipc_close_handle(p)
        self._peer_opened = []
        if getattr(self, "_native", None):
            lib.mfhn_dist_destroy(self._native)
            self._native = None

    # -- peer-memory mode: boundary cells access the owners' vectors over NVLink ---------------------
    def enable_peer(self):
        """Allocates an exportable (src, dst) vector pair, exchanges the CUDA IPC handles and registers
        the peers' vectors with the native operator.  vmult on exactly this pair then runs without
        pack / unpack and without a data-path collective (mfhn_dist_vmult_peer)."""
        import ctypes as C

        torch, dist, part = self.torch, self.dist, self.part
        if self._native is None:
            raise capi.MfhnError(1, "peer mode needs the native (NCCL) operator")
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        n = self.n_owned + self.n_ghost
        self._peer_src_vec = DeviceVector(n, self.dtype, self.device, padding=capi.VECTOR_PADDING)
        self._peer_dst_vec = DeviceVector(n, self.dtype, self.device, padding=capi.VECTOR_PADDING)
        mine = torch.cat([self._peer_src_vec.ipc_handle(), self._peer_dst_vec.ipc_handle()]).to(self.device)
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=self.group)
        psrc = (C.c_void_p * world)()
        pdst = (C.c_void_p * world)()
        self._peer_opened = []
        needed = set(int(o) for o in np.unique(part.ghost_owner)) if part.n_ghost else set()
        for r in range(world):
            if r == rank or r not in needed:
                continue
            hcpu = allh[r].cpu().contiguous()
            a, b = C.c_void_p(), C.c_void_p()
            check(lib.mfhn_ipc_open_handle(hcpu[:64].contiguous().data_ptr(), C.byref(a)))
            check(lib.mfhn_ipc_open_handle(hcpu[64:].contiguous().data_ptr(), C.byref(b)))
            psrc[r], pdst[r] = a.value, b.value
            self._peer_opened += [a.value, b.value]
        owner = np.ascontiguousarray(part.ghost_owner, dtype=np.int32)
        remote = np.ascontiguousarray(part.ghost_global - part.rank_begin[part.ghost_owner], dtype=np.int64) if part.n_ghost else np.zeros(0, np.int64)
        check(lib.mfhn_dist_enable_peer(self._native, self._peer_src_vec.ptr, self._peer_dst_vec.ptr, psrc, pdst,
                                        owner.ctypes.data_as(C.c_void_p), remote.ctypes.data_as(C.c_void_p)))
        # barriers as flag exchanges in each other's memory (MFHN_PEER_BARRIER=nccl keeps the 4-byte all-reduces)
        import os

        self.peer_barrier = os.environ.get("MFHN_PEER_BARRIER", "flags")
        if self.peer_barrier == "flags":
            self._peer_flags_vec = DeviceVector(world + 4, torch.float32, self.device)  # (world + 1) x uint32, zeroed
            mine = self._peer_flags_vec.ipc_handle().to(self.device)
            allf = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allf, mine, group=self.group)
            pflags = (C.c_void_p * world)()
            for r in range(world):
                if r == rank:
                    continue
                a = C.c_void_p()
                check(lib.mfhn_ipc_open_handle(allf[r].cpu().contiguous().data_ptr(), C.byref(a)))
                pflags[r] = a.value
                self._peer_opened.append(a.value)
            check(lib.mfhn_dist_enable_peer_flags(self._native, self._peer_flags_vec.ptr, pflags))
        dist.barrier(group=self.group)  # every rank has opened what it needs before anyone proceeds
        self.peer_src, self.peer_dst = self._peer_src_vec.tensor, self._peer_dst_vec.tensor
        return self.peer_dst, self.peer_src

    # -- building blocks ---------------------------------------------------------
    def _global_rank(self, p):
        return p if self.group is None else self.dist.get_global_rank(self.group, p)

    def _pack(self, vec):
        for p in self.import_peers:
            idx, buf = self.import_idx[p], self.send_buf[p]
            if self.cuda:
                stream = self.torch.cuda.current_stream(self.device).cuda_stream
                check(lib.mfhn_pack(self.number, buf.data_ptr(), vec.data_ptr(), idx.data_ptr(), idx.numel(), stream))
                self.n_launches += 1
            else:
                buf.copy_(vec[idx.long()])

    def _unpack_add(self, vec):
        for p in self.import_peers:
            idx, buf = self.import_idx[p], self.recv_buf[p]
            if self.cuda:
                stream = self.torch.cuda.current_stream(self.device).cuda_stream
                check(lib.mfhn_unpack_add(self.number, vec.data_ptr(), buf.data_ptr(), idx.data_ptr(), idx.numel(), stream))
                self.n_launches += 1
            else:
                vec.index_add_(0, idx.long(), buf)

    def _exchange(self, sends, recvs):
        """sends / recvs: lists of (tensor, peer).  Grouped non-blocking p2p."""
        dist = self.dist
        ops = [dist.P2POp(dist.irecv, t, self._global_rank(p), group=self.group) for t, p in recvs]
        ops += [dist.P2POp(dist.isend, t, self._global_rank(p), group=self.group) for t, p in sends]
        if not ops:
            return
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def _ghost_view(self, vec, p):
        a, b = self.part.ghost_ranges[p]
        return vec[self.n_owned + a:self.n_owned + b]

    # -- LinearAlgebra::distributed::Vector interface --------------------------------
    def update_ghost_values(self, vec):
        self._pack(vec)
        self._exchange([(self.send_buf[p], p) for p in self.import_peers], [(self._ghost_view(vec, p), p) for p in self.ghost_peers])

    def compress_add(self, vec):
        self._exchange([(self._ghost_view(vec, p), p) for p in self.ghost_peers], [(self.recv_buf[p], p) for p in self.import_peers])
        self._unpack_add(vec)
        vec[self.n_owned:].zero_()

    def zero_out_ghost_values(self, vec):
        vec[self.n_owned:].zero_()

    # -- the overlapped cell loop -------------------------------------------------------
    def vmult(self, op, dst, src, zero_dst=False):
        torch = self.torch
        s0, s1, s2, s3 = self.seg
        if zero_dst:
            dst.zero_()
        if not self.cuda:
            self.update_ghost_values(src)
            self.local_apply(dst, src, s0, s3)
            self.compress_add(dst)
            return
        if self._native is not None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if getattr(self, "peer_src", None) is not None and src.data_ptr() == self.peer_src.data_ptr() and dst.data_ptr() == self.peer_dst.data_ptr():
                check(lib.mfhn_dist_vmult_peer(self._native, stream, 0))  # the registered pair: peer-memory path
            else:
                check(lib.mfhn_dist_vmult(self._native, dst.data_ptr(), src.data_ptr(), stream, 0))
            return
        main = torch.cuda.current_stream(self.device)
        comm = self.comm_stream
        self._pack(src)
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            self._exchange([(self.send_buf[p], p) for p in self.import_peers], [(self._ghost_view(src, p), p) for p in self.ghost_peers])
        if s1 > s0:
            self.local_apply(dst, src, s0, s1)  # interior A overlaps the import
            self.n_launches += 1
        main.wait_stream(comm)
        if s3 > s2:
            self.local_apply(dst, src, s2, s3)  # boundary cells need the ghosts
            self.n_launches += 1
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            self._exchange([(self._ghost_view(dst, p), p) for p in self.ghost_peers], [(self.recv_buf[p], p) for p in self.import_peers])
        if s2 > s1:
            self.local_apply(dst, src, s1, s2)  # interior B overlaps the compress
            self.n_launches += 1
        main.wait_stream(comm)
        self._unpack_add(dst)
        dst[self.n_owned:].zero_()

    # -- CUDA graph of one vmult on fixed vectors (the benchmark loop of benchmark_03.h:475-499
    #    applies the operator to the same src / dst 100 times): removes the host launch latency
    #    of the ~20 small launches (pack, 3 cell partitions, NCCL groups, unpack) per vmult
    def capture(self, op, dst, src):
        torch = self.torch
        for _ in range(3):  # warm up NCCL connections and lazy initialisations outside the capture
            self.vmult(op, dst, src)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.vmult(op, dst, src)
        return graph

    def launches_per_vmult(self, cell_launches=None):
        """Kernels per vmult: cell kernels (measured by the caller through LaplaceOperator.launch_count, or one
        per non-empty partition) plus the pack / unpack kernels of the exchange."""
        s0, s1, s2, s3 = self.seg
        cells = int(s1 > s0) + int(s2 > s1) + int(s3 > s2) if cell_launches is None else int(cell_launches)
        if self._native is not None and getattr(self, "peer_src", None) is not None:
            return cells + (2 if getattr(self, "peer_barrier", "nccl") == "flags" else 0)  # two barrier kernels
        if self._native is not None:
            return cells + 2 * int(self.part.n_import_indices() > 0)  # one pack + one unpack kernel
        return cells + 2 * len(self.import_peers)
