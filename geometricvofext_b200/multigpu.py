"""Decomposed (one process per GPU) SimPLIC step.

The reference runs in parallel the OpenFOAM way: one MPI rank per scotch sub-domain, processor
patches, ~10 halo swaps / reductions per step.  Here every rank owns a box of
cells of the global mesh and carries a halo of `G` cell layers around it (overlapping
decomposition): the whole reconstruct()+advect() sequence runs on the extended block with the
unchanged single-GPU kernels, and ONE exchange per step refreshes the halo alpha from the
owning ranks.  Dependency radius of one step: LS normal (1 point layer) -> plane/flux of the
upwind cell (1 face layer) -> bounding corrections (1 layer per sweep that actually moves
fluid), so with G = 4 the owned cells reproduce the single-domain result (to round-off of the
bounding order across the cut, which the reference itself does not preserve in parallel).

Transport is torch.distributed: NCCL over NVLink/NVSwitch between GPUs (device tensors wrapped
around the solver's own alpha buffer -- no host staging), gloo on CPU for the tests.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

from . import capi, fields
from .mesh import hex_block
from .solver import SolveVofEqu

HALO = 4


def block_grid(p):
    """(px, py, pz) with px*py*pz == p, as cubic as possible, z split first (slabs for small p)."""
    best = None
    for px in range(1, p + 1):
        if p % px:
            continue
        for py in range(1, p // px + 1):
            if (p // px) % py:
                continue
            pz = p // px // py
            if px <= py <= pz:
                cand = (pz - px, (px, py, pz))
                if best is None or cand < best:
                    best = cand
    return best[1]


class Decomposition:
    """Box decomposition of an N=(Nx,Ny,Nz) hex mesh over `world` ranks with a G-layer halo."""

    def __init__(self, n, world, halo=HALO, length=(1.0, 1.0, 1.0)):
        self.N = np.array(n if np.ndim(n) else (n, n, n), dtype=np.int64)
        self.world, self.G = world, halo
        self.length = tuple(float(x) for x in length)
        self.grid = np.array(block_grid(world), dtype=np.int64)

    def coords(self, rank):
        px, py, pz = self.grid
        return np.array([rank % px, (rank // px) % py, rank // (px * py)], dtype=np.int64)

    def owned_box(self, rank):
        c = self.coords(rank)
        lo = (self.N * c) // self.grid
        hi = (self.N * (c + 1)) // self.grid
        return lo, hi

    def ext_box(self, rank):
        lo, hi = self.owned_box(rank)
        return np.maximum(lo - self.G, 0), np.minimum(hi + self.G, self.N)

    @staticmethod
    def _overlap(a, b):
        lo, hi = np.maximum(a[0], b[0]), np.minimum(a[1], b[1])
        return (lo, hi) if np.all(hi > lo) else None

    @staticmethod
    def _local_ids(box, ext):
        """local cell ids (natural order of the extended block) of the cells of `box`, in natural order"""
        lo, hi = box
        elo, ehi = ext
        nx, ny = (ehi - elo)[0], (ehi - elo)[1]
        k, j, i = np.meshgrid(np.arange(lo[2], hi[2]), np.arange(lo[1], hi[1]), np.arange(lo[0], hi[0]), indexing="ij")
        return ((i - elo[0]) + nx * ((j - elo[1]) + ny * (k - elo[2]))).reshape(-1).astype(np.int64)

    def plan(self, rank):
        """owned ids + per-neighbour send/recv id lists (both sides enumerate the same global order)."""
        ext = self.ext_box(rank)
        own = self.owned_box(rank)
        plan = {"owned": self._local_ids(own, ext), "send": {}, "recv": {}}
        for r in range(self.world):
            if r == rank:
                continue
            ov = self._overlap(ext, self.owned_box(r))          # my halo cells owned by r
            if ov is not None:
                plan["recv"][r] = self._local_ids(ov, ext)
            ov = self._overlap(own, self.ext_box(r))              # my owned cells inside r's halo
            if ov is not None:
                plan["send"][r] = self._local_ids(ov, ext)
        return plan

    def rank_mesh(self, rank):
        lo, hi = self.ext_box(rank)
        return hex_block(self.N, lo=lo, hi=hi, length=self.length, cut_as_wall=True)


class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DecomposedSolveVofEqu:
    """solveVofEqu on one rank of a decomposed hex mesh; same member names as SolveVofEqu."""

    def __init__(self, n, controls, rank, world, lib=None, device=None, halo=HALO, length=(1.0, 1.0, 1.0)):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = rank, world
        self.dec = Decomposition(n, world, halo, length)
        self.mesh = self.dec.rank_mesh(rank)
        self.plan = self.dec.plan(rank)
        self.on_gpu = device is not None
        comm = (rank, 1, device if device is not None else -1)   # each handle is a single-domain solver of its block
        self.s = SolveVofEqu(self.mesh, controls, lib=lib, comm=comm)
        self.device = device
        self.owned = self.plan["owned"]
        if self.on_gpu:
            import torch
            self.torch = torch
            dev = torch.device("cuda", device)
            # one gather and one scatter per step whatever the number of neighbours: the send indices of all peers are
            # concatenated (each peer's message is a slice of the packed buffer), and so are the receive indices
            def cat(d):
                ranks = sorted(d)
                sizes = [len(d[r]) for r in ranks]
                idx = np.concatenate([np.asarray(d[r], dtype=np.int64) for r in ranks]) if ranks else np.zeros(0, np.int64)
                return ranks, sizes, idx
            self.send_ranks, self.send_sizes, sidx = cat(self.plan["send"])
            self.recv_ranks, self.recv_sizes, ridx = cat(self.plan["recv"])
            self.send_idx_all = torch.as_tensor(sidx, device=dev)
            self.recv_idx_all = torch.as_tensor(ridx.astype(np.int32), device=dev)
            self.send_buf_all = torch.empty(len(sidx), dtype=torch.float64, device=dev)
            self.recv_buf_all = torch.empty(len(ridx), dtype=torch.float64, device=dev)
            st = C.c_void_p()
            self.s._chk(self.s.lib.svof_get_stream(self.s._h, C.byref(st)))
            self.ext_stream = torch.cuda.ExternalStream(st.value, device=dev)
        self.halo_bytes = 8 * sum(len(v) for v in self.plan["recv"].values())

    # -- halo exchange: alpha of halo cells <- owning rank ---------------------------------------
    def exchange_alpha(self):
        dist = self.dist
        if self.world == 1:
            return
        if self.on_gpu:
            # Fully stream-ordered: gather -> NCCL send/recv -> scatter are enqueued on the solver's own CUDA
            # stream (wrapped as a torch ExternalStream), so the step needs no host synchronisation.
            torch = self.torch
            p = C.c_void_p()
            self.s._chk(self.s.lib.svof_device_ptr(self.s._h, capi.F_ALPHA, C.byref(p)))
            with torch.cuda.stream(self.ext_stream):
                a = torch.as_tensor(_DevArr(p.value, self.s.nC), device=torch.device("cuda", self.device))
                torch.index_select(a, 0, self.send_idx_all, out=self.send_buf_all)
                ops, o = [], 0
                for r, n in zip(self.send_ranks, self.send_sizes):
                    ops.append(dist.P2POp(dist.isend, self.send_buf_all[o:o + n], r))
                    o += n
                o = 0
                for r, n in zip(self.recv_ranks, self.recv_sizes):
                    ops.append(dist.P2POp(dist.irecv, self.recv_buf_all[o:o + n], r))
                    o += n
                for w in dist.batch_isend_irecv(ops):
                    w.wait()          # stream-level wait (no host block) for NCCL work
            # halo cells <- received values, mixed-cell bitmap and patch values kept up to date (same stream)
            self.s._chk(self.s.lib.svof_scatter_alpha_device(self.s._h, self.recv_idx_all.data_ptr(), self.recv_buf_all.data_ptr(),
                                                             self.recv_idx_all.numel()))
        else:
            import torch
            a = self.s.alpha()
            reqs, bufs = [], {}
            for r, idx in self.plan["send"].items():
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[idx])), r))
            for r, idx in self.plan["recv"].items():
                bufs[r] = torch.empty(len(idx), dtype=torch.float64)
                reqs.append(dist.irecv(bufs[r], r))
            for q in reqs:
                q.wait()
            for r, idx in self.plan["recv"].items():
                a[idx] = bufs[r].numpy()
            self.s.setAlpha(a)

    # -- the reference's member functions -----------------------------------------------------------
    def setAlpha(self, alpha_local):
        self.s.setAlpha(alpha_local)

    def setPhi(self, phi):
        self.s.setPhi(phi)

    def setU(self, U, Ub=None):
        self.s.setU(U, Ub)

    def reconstruct(self):
        self.s.reconstruct()

    def advect(self, dt, Sp=None, Su=None):
        self.s.advect(dt, Sp, Su)
        self.exchange_alpha()

    def step(self, dt):
        """reconstruct() + advect(dt) as one CUDA-graph launch on the extended block, then the halo swap."""
        self.s.step(dt)
        self.exchange_alpha()

    def alpha_owned(self):
        return self.s.alpha()[self.owned]

    def owned_global_ids(self):
        return self.mesh.cell_global[self.owned]

    def volume(self):
        """gSum(alpha*V) over the owned cells of all ranks."""
        import torch
        v = float(np.sum(self.s.alpha()[self.owned] * self.s.field(capi.F_V)[self.owned]))
        t = torch.tensor([v], dtype=torch.float64)
        if self.on_gpu:
            t = t.cuda(self.device)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def close(self):
        self.s.close()


# ------------------------------------------------------------------------------------ bench ----
def bench(args, controls, metric, unit):
    """N-GPU leg of bench.py: one rank per GPU, each owning a 256^3 block of the global mesh."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL's version banner goes to stdout: keep rank 0's stdout to the one JSON line
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    grid = np.array(block_grid(world))
    # weak scaling: the LeVeque problem tiled grid[0] x grid[1] x grid[2] times -- every GPU owns one unit cube
    # (args.n^3 cubic cells, its own sphere; the velocity field is 1-periodic and tangential on the tile faces),
    # so the per-GPU work is identical to the N=1 run and the cells stay cubic (the analytic face fluxes are
    # discretely divergence free only on cubic cells).  8 GPUs = 512^3 cells in total.
    n_global = (args.n * grid).tolist()
    t0 = time.perf_counter()
    ds = DecomposedSolveVofEqu(n_global, controls, rank, world, device=local, length=grid.astype(float).tolist())
    s = ds.s
    tile = ds.dec.coords(rank).astype(float)
    a0 = fields.sphere_alpha_quadrature(ds.mesh, centre=tuple(tile + 0.35))
    dt = 0.2 / args.n
    C_, Cf, Sf = s.field(capi.F_C), s.field(capi.F_CF), s.field(capi.F_SF)
    f = fields.u_factor(dt, dt, 6.0)
    U = fields.leveque_velocity(C_) * f
    phi = fields.face_flux(Cf, Sf) * f
    s.setAlpha(a0)
    s.setPhi(phi)
    s.setU(U, np.zeros((s.nBF, 3)))
    ds.exchange_alpha()
    setup_s = time.perf_counter() - t0

    def step():
        ds.step(dt)

    for _ in range(max(3, args.warmup)):
        step()
    s.synchronize()
    l0 = s.info(capi.I_GPU_LAUNCHES)
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.lib.svof_mark(s._h, 0)
    ev0.record()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        step()
    s.lib.svof_mark(s._h, 1)
    ev1.record()
    s.synchronize()
    torch.cuda.synchronize()
    dist.barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    ms = C.c_double()
    s.lib.svof_elapsed_ms(s._h, 0, 1, C.byref(ms))
    # device time of the region on this rank = max(own-stream events, torch-stream events); job time = max over ranks
    t = torch.tensor([max(ms.value, ev0.elapsed_time(ev1))], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    cells = torch.tensor([float(len(ds.owned))], dtype=torch.float64, device="cuda")
    dist.all_reduce(cells)
    launches = int(s.info(capi.I_GPU_LAUNCHES) - l0)
    value = float(cells.item()) * args.steps / (total_ms * 1e-3)
    vol = ds.volume()
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "LeVeque 3-D deformation tiled %dx%dx%d: one unit cube (%d^3 cubic hex cells, sphere r=0.15) per GPU, "
                                   "%dx%dx%d cells in total (8 GPUs: 512^3 cells, the size of BASELINE.json configs[4])" %
                                   (grid[0], grid[1], grid[2], args.n, n_global[0], n_global[1], n_global[2]),
                       "cells": int(cells.item()), "dt": dt, "controls": controls, "halo_layers": ds.dec.G,
                       "halo_bytes_per_step_rank0": ds.halo_bytes, "exchange": "1 alpha halo swap per step, NCCL send/recv",
                       "l2": "inputs larger than L2", "timing": "CUDA events, max over ranks; host wall %.3f ms/step" % (wall_ms / args.steps),
                       "setup_s": setup_s, "volume": vol},
            "gpu_launches": launches,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "multi-GPU leg is device resident; the host-buffer end-to-end number is measured at N=1"},
        }
        print(json.dumps(line))
    dist.barrier()
    if rank == 0 or os.environ.get("SVOF_PROFILE_ALL"):
        ds.close()
    dist.barrier()
    dist.destroy_process_group()
