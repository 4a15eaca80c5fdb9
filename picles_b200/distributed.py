"""y-strip decomposition of the particle-in-cell step: one process per GPU, one strip per process.

The reference has no domain decomposition: particles are independent
during the ODE phase and couple only through the bilinear deposit, whose reach is a few
rows.  Each rank owns a contiguous block of rows of every per-node plane; between the
advance and the projection gather it ships the deposit records (5 f64 + packed cell) of
its first / last `halo` rows to the two y-neighbours.  Every rank then sums its own
nodes in the reference's canonical order, so the fields are bit-identical for any rank
count.  There is no other data-path collective.

Two transports for the exchange:

  "nccl-lib"   picles_halo_exchange: pack, ncclSend/ncclRecv (one group) and unpack are all
               enqueued on the handle's stream inside libpicles_b200.so.  torch.distributed
               is used once, to broadcast the 128-byte NCCL id.  This is the GPU path.
  "torch-p2p"  torch.distributed.batch_isend_irecv on tensors that alias the engine's halo
               buffers.  With the gloo backend and a host build of the device code this is
               what the CPU tests run (world_size 2); with NCCL it is an alternative for
               hosts that want every collective to stay in torch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def strip_bounds(Ny: int, world: int):
    """[(j0, j1)) rows of every rank: contiguous, balanced to one row."""
    return [(Ny * r // world, Ny * (r + 1) // world) for r in range(world)]


def strip_bounds_weighted(row_cost, world: int, min_rows: int = 1):
    """[(j0, j1)) rows of every rank for rows of unequal cost (land, pole rows): contiguous strips
    whose summed cost is as even as a prefix-sum split allows.  row_cost: one non-negative weight
    per global row; every strip keeps at least `min_rows` rows (>= the halo width)."""
    c = np.maximum(np.asarray(row_cost, np.float64), 0.0)
    Ny = c.size
    if world * min_rows > Ny:
        raise ValueError(f"{Ny} rows cannot hold {world} strips of at least {min_rows} rows")
    cum = np.concatenate([[0.0], np.cumsum(c)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        j = int(np.searchsorted(cum, target))
        if j > 0 and abs(cum[j - 1] - target) <= abs(cum[min(j, Ny)] - target):
            j -= 1
        j = max(j, cuts[-1] + min_rows)
        j = min(j, Ny - (world - r) * min_rows)
        cuts.append(j)
    cuts.append(Ny)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def row_cost_model(mask, periodic_boundary: bool, reach_rows=None, gather_weight: float = 0.068):
    """cost of every global row in particle units, for strip_bounds_weighted: the active particles
    of the row (the advance is ~90-98 % of a step) plus the gather/remesh of all its nodes.  The gather
    visits a window of (2R+1)^2 candidate sources per node, R = how many cells the deposits of the rows
    around reach (reach_rows: per-row estimate, e.g. from the small cells near a tripolar pole), so its
    cost per node is gather_weight * (2R+1)^2 / 9.  Weights from the measured kernel times of the
    tripolar + land configuration (profiles/README.md, round 2): advance 1.12 ns per active particle,
    gather + remesh 0.076 ns per node at reach 1 (0.068 particle units), 0.39 ns per node over the rows
    near the pole (reach 3: 49/9 x 0.068 = 0.37)."""
    m = np.asarray(mask)
    active = (m == 1) | ((m == 3) if periodic_boundary else False)
    R = np.ones(m.shape[0]) if reach_rows is None else np.maximum(np.asarray(reach_rows, np.float64), 1.0)
    return active.sum(axis=1).astype(np.float64) + gather_weight * m.shape[1] * (2.0 * R + 1.0) ** 2 / 9.0


def row_cost_measured(active_rows, nodes_per_row, reach_rows, bounds, ms_advance, ms_gather, split=False):
    """cost of every global row from a calibration run on the strips `bounds`: each strip's measured advance
    time spread over its rows by their active particles (captures regional differences in substeps), its measured
    gather + remesh time spread over its rows by the window area (2R+1)^2 of the rows' measured reach.
    active_rows, reach_rows: one value per global row; ms_advance, ms_gather: one value per strip.
    split=True returns the two parts (advance, gather) instead of their sum."""
    a = np.asarray(active_rows, np.float64)
    w = nodes_per_row * (2.0 * np.maximum(np.asarray(reach_rows, np.float64), 1.0) + 1.0) ** 2
    adv, gat = np.zeros_like(a), np.zeros_like(a)
    for (j0, j1), ta, tg in zip(bounds, ms_advance, ms_gather):
        adv[j0:j1] = ta * a[j0:j1] / max(a[j0:j1].sum(), 1.0)
        gat[j0:j1] = tg * w[j0:j1] / max(w[j0:j1].sum(), 1.0)
    return (adv, gat) if split else adv + gat


def overlapped_step_estimate(adv_cost, gather_cost, bounds):
    """what picles_step_strip's overlapped step costs on the slowest strip, up to a constant: every strip gathers only
    after the all-reduce of the reach, which completes when the SLOWEST strip's advance does (the exchange's kernels
    find no free SM beside a persistent interior launch), so a strip's step is max_r(advance_r) + its own gather —
    measured on 2 and on 8 GPUs with a residual of 0.11-0.14 ms that is the same on every strip
    (profiles/r02_bench_n2.json, r02_bench_n8.json)."""
    A = max(float(np.sum(adv_cost[j0:j1])) for j0, j1 in bounds)
    G = max(float(np.sum(gather_cost[j0:j1])) for j0, j1 in bounds)
    return A + G


def strip_bounds_overlapped(adv_cost, gather_cost, world: int, min_rows: int = 1, blends=(0.0, 0.25, 0.5, 0.75, 1.0)):
    """strips for the overlapped step: the prefix-sum split of advance + λ·gather, λ in `blends`, that minimises
    overlapped_step_estimate — in practice the split that evens out the ADVANCE (λ = 0) unless that piles too much
    gather on one strip; λ = 1 is strip_bounds_weighted on the summed cost.  Returns (bounds, λ, estimate)."""
    adv = np.maximum(np.asarray(adv_cost, np.float64), 0.0)
    gat = np.maximum(np.asarray(gather_cost, np.float64), 0.0)
    best = None
    for lam in blends:
        b = strip_bounds_weighted(adv + lam * gat, world, min_rows)
        est = overlapped_step_estimate(adv, gat, b)
        if best is None or est < best[2] - 1e-12:
            best = (b, float(lam), est)
    return best


def neighbours(rank: int, world: int, periodic_y: bool):
    """(lo, hi): ranks owning the rows below / above this strip, -1 at a domain edge."""
    if world == 1:
        return -1, -1
    lo, hi = rank - 1, rank + 1
    if periodic_y:
        return lo % world, hi % world
    return (lo if lo >= 0 else -1), (hi if hi < world else -1)


def slab(a, j0, j1, lead=0):
    """rows [j0, j1) of a global (…, Ny, Nx) array (None passes through)."""
    if a is None:
        return None
    a = np.asarray(a)
    return a[(slice(None),) * lead + (slice(j0, j1),)]


class _DeviceBytes:
    """zero-copy view of `nbytes` of device memory for torch.as_tensor"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class NcclLibTransport:
    """halo exchange inside the library (picles_comm_init / picles_halo_exchange)"""

    name = "nccl-lib"

    def __init__(self, eng, rank, world, nccl_path=None, bcast=None):
        lib = eng.lib
        path = nccl_path.encode() if nccl_path else None
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            eng._check(lib.picles_comm_unique_id(idbuf, path), None)
        payload = [idbuf.raw if rank == 0 else None]
        if bcast is None:
            import torch.distributed as dist
            dist.broadcast_object_list(payload, src=0)
        else:
            payload[0] = bcast(payload[0])
        eng._check(lib.picles_comm_init(eng.h, payload[0], int(rank), int(world), path))
        self.eng = eng

    def exchange(self, lo, hi):
        e = self.eng
        e._check(e.lib.picles_halo_exchange(e.h, int(lo), int(hi)))


class TorchP2PTransport:
    """halo exchange through torch.distributed point-to-point operations"""

    name = "torch-p2p"

    def __init__(self, eng, rank, world):
        import torch
        self.torch = torch
        self.eng = eng
        self._views()

    def _views(self):
        torch, eng = self.torch, self.eng
        bufs, nb = eng.halo_buffers()
        self.nb = nb
        self.on_device = not isinstance(bufs[0], np.ndarray)
        if self.on_device:
            dev = torch.device("cuda", eng.device)
            self.t = [torch.as_tensor(_DeviceBytes(p, nb), device=dev) for p in bufs]
        else:
            self.t = [torch.from_numpy(b) for b in bufs]

    def validate_reach(self):
        """the halo width is validated every step by a max-all-reduce of the strips' reach;
        deposits that reach further than the rows exchanged widen the exchange before it happens"""
        import torch.distributed as dist
        e = self.eng
        if not hasattr(e, "halo_widen"):
            return
        r = self.torch.tensor([e.reach()], dtype=self.torch.int32, device=f"cuda:{e.device}" if self.on_device else "cpu")
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
        need = int(r.item())
        e.set_global_reach(need)
        if need > e.halo_rows()[0]:
            e.halo_widen(need)
            self._views()

    def exchange(self, lo, hi):
        import torch.distributed as dist
        e = self.eng
        self.validate_reach()
        e.halo_pack()
        send_lo, send_hi, recv_lo, recv_hi = self.t
        ops = []
        # tag 1: rows travelling down (my first rows -> lower neighbour's upper halo),
        # tag 0: rows travelling up.  Issue order sends lo,hi / receives hi,lo so a two-strip
        # ring (lo == hi) pairs correctly on backends that ignore tags (NCCL).
        if lo >= 0:
            ops.append(dist.P2POp(dist.isend, send_lo, lo, tag=1))
        if hi >= 0:
            ops.append(dist.P2POp(dist.isend, send_hi, hi, tag=0))
        if hi >= 0:
            ops.append(dist.P2POp(dist.irecv, recv_hi, hi, tag=1))
        if lo >= 0:
            ops.append(dist.P2POp(dist.irecv, recv_lo, lo, tag=0))
        if ops:
            if self.on_device:
                e.synchronize()  # pack ran on the handle's stream; NCCL runs on torch's
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            if self.on_device:
                self.torch.cuda.current_stream().synchronize()
        e.halo_unpack()


class StripStepper:
    """time_step! of one y-strip: upload winds, advance, halo exchange, gather, remesh."""

    def __init__(self, eng, rank, world, periodic_y=False, transport="nccl-lib", **kw):
        if world > 1 and eng.halo < 1:
            raise ValueError("a strip decomposition needs halo >= 1")
        self.eng, self.rank, self.world = eng, int(rank), int(world)
        self.lo, self.hi = neighbours(self.rank, self.world, periodic_y)
        if world == 1:
            self.transport = None
        elif transport == "nccl-lib":
            self.transport = NcclLibTransport(eng, rank, world, **kw)
        elif transport == "torch-p2p":
            self.transport = TorchP2PTransport(eng, rank, world)
        else:
            raise ValueError(f"unknown transport {transport!r}")

    def step(self, t, DT, host_ptrs=None, winds=None):
        """host_ptrs = (u_t1, v_t1) raw host pointers of this strip's wind at t+DT (the
        previous t+DT level becomes the t level); winds = (u_t, v_t, u_t1, v_t1) arrays;
        neither: reuse what is on the device."""
        e = self.eng
        if isinstance(self.transport, NcclLibTransport):
            # one C-ABI call: wind upload pipelined against the advance, exchange on the stream
            if host_ptrs is not None:
                e.step_strip_raw(t, DT, None, None, host_ptrs[0], host_ptrs[1], self.lo, self.hi)
            else:
                e.step_strip(t, DT, *(winds if winds is not None else (None,) * 4), lo=self.lo, hi=self.hi)
            return
        if host_ptrs is not None:
            e.upload_winds_raw(None, None, host_ptrs[0], host_ptrs[1])
        elif winds is not None:
            e.upload_winds(*winds)
        e.step_advance(t, DT)
        if self.transport is not None:
            self.transport.exchange(self.lo, self.hi)
        e.step_project_remesh(t, DT)

    def step_wind_mesh(self, t, DT, n_mid=0):
        """one strip step with every wind level sampled from the engine's resident wind mesh
        (picles_set_wind_mesh): one C-ABI call with the in-library exchange, phase-split otherwise"""
        e = self.eng
        if self.transport is None or isinstance(self.transport, NcclLibTransport):
            e.step_wind_mesh(t, DT, n_mid, self.lo, self.hi)
            return
        if hasattr(e, "stage_wind_mesh"):      # B200Engine: levels sampled on the device, nothing uploaded
            e.stage_wind_mesh(t, DT, n_mid)
        else:                                   # host build of the device code (CPU tests)
            if n_mid:
                lv = [e.sample_wind_mesh(t + DT * float(k) / float(n_mid + 1)) for k in range(1, n_mid + 1)]
                e.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
            e.upload_winds(*e.sample_wind_mesh(t), *e.sample_wind_mesh(t + DT))
        e.step_advance(t, DT)
        self.transport.exchange(self.lo, self.hi)
        e.step_project_remesh(t, DT)
