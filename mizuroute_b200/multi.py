"""Multi-domain routing: tributary domains on every rank, mainstem on rank 0 after a downstream-only hand-off.

This is the reference's `mpi_route` (mpi_process.f90:1088-1342) with the MPI hub replaced by NCCL point-to-point:

    reference                                                     here
    ---------------------------------------------------------     -----------------------------------------------
    main_route(tributary slab) on every rank        (:1217)       DomainSet.trib.route_resident / route_batch
    mpi_comm_river_flux  gather REACH_Q             (:1245)  \\
    mpi_comm_flux        gather BASIN_QR            (:1256)   >   one record per outlet and step in the export
    mpi_comm_kwt_state   gather wave particles      (:1265)  /    buffer -> `exchange_rows` (isend/irecv group)
    main_route(mainstem slab + outlet ghosts) on rank 0 (:1294)   DomainSet.main.route_resident
    mpi_comm_kwt_state   scatter stripped waves back (:1319)      not needed (owners strip their own particles)

`exchange_rows` is backend-agnostic (`torch.distributed`: nccl with CUDA tensors, gloo with CPU tensors in the
CPU tests).  `route_decomposed_local` runs all domains of a decomposition inside ONE process on one GPU
(device-to-device hand-off), which is how the parity tests check the decomposition against a single-domain run.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import capi, partition
from .network import RiverNetwork, RouteOptions, RouteParams
from .route import Router


def record_len(opts: RouteOptions) -> int:
    """doubles per (outlet, step) record: REACH_Q per route | BASIN_QR(1) | numWaves | numRouted | QF[24] | TR[24]"""
    return len(opts.route_opt) + 3 + 2 * capi.KW_PITCH


def exchange_rows(local_rows, gathered_rows, dec: partition.Decomposition, rank: int, world: int):
    """Move every rank's rows (one row per tributary outlet it owns, `dec.outlets_of(rank)` order) into rank 0's
    `gathered_rows` (one row per outlet, `dec.outlets` order).  Tensors are 2-D [rows, row_len] torch tensors on
    the backend's device; `gathered_rows` is only used on rank 0.  One isend/irecv group = one ncclGroup."""
    import torch.distributed as dist
    ops = []
    if rank == 0:
        lo, hi = dec.slot_range(0)
        if hi > lo:
            gathered_rows[lo:hi].copy_(local_rows[: hi - lo])
        for r in range(1, world):
            lo, hi = dec.slot_range(r)
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, gathered_rows[lo:hi], r))
    else:
        lo, hi = dec.slot_range(rank)
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, local_rows[: hi - lo], 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class DomainSet:
    """The domains one rank routes: its tributary forest and, on rank 0, the mainstem with outlet ghosts."""

    def __init__(self, net: RiverNetwork, params: RouteParams, opts: RouteOptions, max_batch: int, rank: int, world: int,
                 device: int = 0, distributed: bool = True, dec: Optional[partition.Decomposition] = None):
        import torch
        self.torch = torch
        self.rank, self.world, self.opts, self.K = rank, world, opts, int(max_batch)
        self.dec = dec if dec is not None else partition.decompose(net, world)
        d = self.dec
        self.dev = torch.device("cuda", device)
        self.trib: Optional[Router] = None           # a rank may hold no tributary domain (rank 0 of a small network)
        self.trib_net: Optional[RiverNetwork] = None
        if d.trib[rank].size:
            self.trib_net = partition.subnetwork(net, d.trib[rank])
            self.trib = Router(self.trib_net, params, opts, device=device, max_batch=max_batch)
        self.my_outlets = d.outlets_of(rank)
        self.rec = record_len(opts)
        self.main: Optional[Router] = None
        self.main_net: Optional[RiverNetwork] = None
        self.exp_t = self.imp_t = None
        self._copy_done = None
        has_main = d.mainstem.size > 0
        if self.my_outlets.size:
            self.trib.set_export(net.segId[self.my_outlets])
            self.exp_t = torch.empty((self.my_outlets.size, self.K * self.rec), dtype=torch.float64, device=self.dev)
            self.trib.set_exchange_buffer(0, self.exp_t.data_ptr(), self.exp_t.numel() * 8)
        if not has_main:
            return
        # ghost parameters: what the owner derived for each outlet (total area, width, count(goodBas))
        mine = np.zeros((self.my_outlets.size, 3))
        if self.my_outlets.size:
            loc = np.searchsorted(d.trib[rank], self.my_outlets)
            mine[:, 0] = self.trib.flux(capi.TOTAREA)[loc]
            mine[:, 1] = self.trib.flux(capi.R_WIDTH)[loc]
            mine[:, 2] = self.trib.flux(capi.NGOOD)[loc]
        if distributed and world > 1:
            g = torch.zeros((d.outlets.size, 3), dtype=torch.float64, device=self.dev)
            exchange_rows(torch.from_numpy(mine).to(self.dev), g, d, rank, world)
            ghost_par = g.cpu().numpy()
        else:
            ghost_par = mine
        self._ghost_par_local = mine
        if rank == 0 and (distributed or world == 1):     # (a local multi-domain driver gathers the parameters itself)
            self._build_mainstem(net, params, opts, device, ghost_par)

    def _build_mainstem(self, net, params, opts, device, ghost_par):
        torch, d = self.torch, self.dec
        self.main_net = partition.mainstem_network(net, d)
        kind = np.where(ghost_par[:, 2] > 0, 2, 1).astype(np.int32)
        self.main = Router(self.main_net, params, opts, device=device, max_batch=self.K,
                           ghosts=(net.segId[d.outlets], kind, ghost_par[:, 0], ghost_par[:, 1]))
        self.imp_t = torch.zeros((d.outlets.size, self.K * self.rec), dtype=torch.float64, device=self.dev)
        self.main.set_exchange_buffer(1, self.imp_t.data_ptr(), self.imp_t.numel() * 8)

    # ------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        if self.trib is not None:
            self.trib.set_stream(cuda_stream)
        if self.main is not None:
            self.main.set_stream(cuda_stream)

    def upload_runoff(self, runoff_global: np.ndarray):
        """runoff_global[K, nHRU_global] (host) -> this rank's HRU columns to its domains."""
        K = runoff_global.shape[0]
        if self.trib is not None:
            self.trib.upload_runoff(np.ascontiguousarray(runoff_global[:, self.trib_net.meta["hru_index"]]))
        if self.main is not None:
            self.main.upload_runoff(np.ascontiguousarray(runoff_global[:, self.main_net.meta["hru_index"]]))
        return K

    def upload_lake_forcing(self, evapo_global: np.ndarray, precip_global: np.ndarray):
        """Lake evaporation / precipitation [K, nHRU_global] of the next routing call, this rank's HRU columns to its domains."""
        for dom, sub in ((self.trib, getattr(self, "trib_net", None)), (self.main, getattr(self, "main_net", None))):
            if dom is not None:
                cols = sub.meta["hru_index"]
                dom.upload_lake_forcing(np.ascontiguousarray(evapo_global[:, cols]), np.ascontiguousarray(precip_global[:, cols]))

    def upload_wm(self, flux_global=None, vol_global=None, vol_jumpstart: bool = False):
        """Water management [K, nRch_global] of the next routing call, this rank's reach columns to its domains (the ghost
        reaches of the mainstem domain receive their reach's values too; they are not routed there)."""
        pick = lambda a, cols: None if a is None else np.ascontiguousarray(a[:, cols])
        for dom, sub in ((self.trib, getattr(self, "trib_net", None)), (self.main, getattr(self, "main_net", None))):
            if dom is not None:
                cols = sub.meta["reach_index"]
                dom.upload_wm(pick(flux_global, cols), pick(vol_global, cols), vol_jumpstart)

    def set_da(self, qmod_option: int = 1, q_blend_period: int = 10, q_err_trend: int = 1):
        """Data assimilation options to every domain of this rank (mr_set_da)."""
        for dom in (self.trib, self.main):
            if dom is not None:
                dom.set_da(qmod_option, q_blend_period, q_err_trend)

    def upload_obs(self, obs_global: np.ndarray, has_record=None):
        """Gauge observations [K, nRch_global] of the next routing call, this rank's reach columns to its domains.  A
        tributary outlet is corrected where it is routed; its ghost in the mainstem domain only hands the corrected flow on."""
        for dom, sub in ((self.trib, getattr(self, "trib_net", None)), (self.main, getattr(self, "main_net", None))):
            if dom is not None:
                dom.upload_obs(np.ascontiguousarray(np.atleast_2d(obs_global)[:, sub.meta["reach_index"]]), has_record)

    def hand_off(self):
        if self.dec.mainstem.size == 0:
            return
        if self.world == 1:
            return
        local = self.exp_t if self.exp_t is not None else self.torch.empty((0, self.K * self.rec), dtype=self.torch.float64, device=self.dev)
        exchange_rows(local, self.imp_t, self.dec, self.rank, self.world)

    def route_resident(self, K: int):
        """K steps of the forcing uploaded last: tributaries everywhere, hand-off, then the mainstem on rank 0."""
        if self.trib is not None:
            self.trib.route_resident(K)
        self.hand_off()
        if self.main is not None:
            self.main.route_resident(K)

    def route_batch_pipelined(self, ro_trib, out_trib, ro_main, out_main, stream_trib, stream_main):
        """End-to-end variant of `route_resident_pipelined`: pinned host forcing in, pinned REACH_Q out, for the
        tributary domain of this rank and (rank 0) the mainstem; uploads, routing, hand-off and downloads of
        consecutive calls overlap (mr_step_batch_async per domain).  Call `wait()` before reading the outputs."""
        self._pipelined(stream_trib, stream_main,
                        (lambda: self.trib.route_batch_async(ro_trib, out_trib)) if self.trib is not None else None,
                        (lambda: self.main.route_batch_async(ro_main, out_main)) if self.main is not None else None)

    def route_resident_pipelined(self, K: int, stream_trib, stream_main):
        """The same, without host waits and with the mainstem on its own stream: the mainstem of this batch overlaps
        the tributaries of the next one (tributaries never depend on the mainstem).  `stream_*` are
        torch.cuda.Stream objects the two routers were bound to with set_stream.  Call `wait()` to collect errors.

        Rank 0 receives on the MAINSTEM stream, so its tributary stream never waits for the other ranks; its own
        outlets are copied export -> import first, and the next tributary batch (which overwrites the export buffer)
        waits for that copy only."""
        self._pipelined(stream_trib, stream_main,
                        (lambda: self.trib.route_resident_async(K)) if self.trib is not None else None,
                        (lambda: self.main.route_resident_async(K)) if self.main is not None else None)

    def _pipelined(self, stream_trib, stream_main, run_trib, run_main):
        torch = self.torch
        with torch.cuda.stream(stream_trib):
            if self._copy_done is not None:
                stream_trib.wait_event(self._copy_done)
            if run_trib is not None:
                run_trib()
            if self.main is None:
                self.hand_off()                            # send: ordered after this rank's tributary kernels
                return
            trib_done = torch.cuda.Event()
            trib_done.record(stream_trib)
        with torch.cuda.stream(stream_main):
            stream_main.wait_event(trib_done)              # (the previous mainstem batch precedes us on this stream)
            lo, hi = self.dec.slot_range(0)
            if hi > lo:
                self.imp_t[lo:hi].copy_(self.exp_t[: hi - lo])
            self._copy_done = torch.cuda.Event()
            self._copy_done.record(stream_main)
            self._recv_others()
            run_main()

    def _recv_others(self):
        import torch.distributed as dist
        if self.world == 1:
            return
        ops = []
        for r in range(1, self.world):
            lo, hi = self.dec.slot_range(r)
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, self.imp_t[lo:hi], r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def wait(self):
        if self.trib is not None:
            self.trib.wait()
        if self.main is not None:
            self.main.wait()

    def launches(self) -> int:
        n = self.trib.info(capi.INFO_LAUNCHES_LAST) if self.trib is not None else 0
        if self.main is not None:
            n += self.main.info(capi.INFO_LAUNCHES_LAST)
        return n


def route_decomposed_local(net: RiverNetwork, params: RouteParams, opts: RouteOptions, runoff: np.ndarray, nparts: int,
                           batch: int, device: int = 0):
    """All domains of an `nparts`-way decomposition in ONE process on one GPU; returns REACH_Q[n_routes, K, nRch] in
    the global reach order and the decomposition.  Used by the parity tests (decomposed == single-domain)."""
    dec = partition.decompose(net, nparts)
    doms: List[DomainSet] = [DomainSet(net, params, opts, batch, r, nparts, device=device, distributed=False, dec=dec)
                             for r in range(nparts)]
    root = doms[0]
    if dec.mainstem.size:
        par = np.concatenate([dm._ghost_par_local for dm in doms], axis=0)
        root._build_mainstem(net, params, opts, device, par)
    K = runoff.shape[0]
    nm = len(opts.route_opt)
    q = np.full((nm, K, net.nRch), np.nan)
    for s in range(0, K, batch):
        ro = np.ascontiguousarray(runoff[s:s + batch])
        k = ro.shape[0]
        for r, dm in enumerate(doms):
            if dm.trib is None:
                continue
            dm.trib.upload_runoff(np.ascontiguousarray(ro[:, dm.trib_net.meta["hru_index"]]))
            dm.trib.route_resident(k)
            q[:, s:s + k, dec.trib[r]] = dm.trib.download_q(k)
            lo, hi = dec.slot_range(r)
            if dec.mainstem.size and hi > lo:
                dm.trib.copy_exchange_to(root.main, 0, lo, hi - lo)
        if root.main is not None:
            root.main.upload_runoff(np.ascontiguousarray(ro[:, root.main_net.meta["hru_index"]]))
            root.main.route_resident(k)
            qm = root.main.download_q(k)
            keep = ~root.main_net.meta["ghost_mask"]
            q[:, s:s + k, root.main_net.meta["reach_index"][keep]] = qm[:, :, keep]
    return q, dec
