"""Sharded statevector over 2/4/8 GPUs (one process per GPU, ``torch.distributed`` / NCCL).

The state of ``n`` qubits is split by its top ``g = log2(world_size)`` PHYSICAL index bits; rank
``r`` owns the amplitudes whose top bits equal ``r``.  Diagonal phases and pair rotations on two
local qubits need no communication.  A rotation touching a sharded qubit is handled by
re-mapping: one all-to-all swaps the ``g`` sharded qubits with ``g`` local ones (each rank keeps
1/P of its shard and sends (P-1)/P of it), after which the engine continues locally.  The
schedule compiler inside ``libxyk.so`` decides where the exchanges go (furthest-next-use rule);
this module only performs the collective it asks for.

torch is used for what it is here for: device buffers handed to NCCL and the process group.
The reference has no multi-device path at all; its selector simply refuses
states that do not fit (backend_selector.py:177-183).
"""

from __future__ import annotations

from ctypes import byref, c_int, c_size_t, c_void_p

import numpy as np

from . import _lib
from .engine import XYKEngine


class _DevicePtr:
    """Minimal ``__cuda_array_interface__`` carrier so torch can wrap an engine buffer without a copy."""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {
            "shape": (n_doubles,),
            "typestr": "<f8",
            "data": (ptr, False),
            "version": 2,
        }


class ShardedXYKEngine:
    """One rank's view of a sharded statevector; every rank must call the same methods in the same order."""

    def __init__(self, n: int, *, device: int, group=None, peer_exchange: bool = True):
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (backend 'nccl') before creating a sharded engine")
        self._torch = torch
        self._dist = dist
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world_size & (self.world_size - 1):
            raise ValueError("world size must be a power of two")
        self.g = self.world_size.bit_length() - 1
        self.n = int(n)
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.eng = XYKEngine(self.n, device=self.device, world_size=self.world_size, rank=self.rank)
        self.n_local = self.eng.n_local
        self.exchanges = 0
        self.exchange_seconds = 0.0
        self.exchange_bytes = 0
        self.exchange_log: list[float] = []  # milliseconds per exchange
        self.barriers = 0
        self._peers_mapped = False
        if peer_exchange:
            self._map_peer_buffers()
            self._peers_mapped = True

    def _map_peer_buffers(self) -> None:
        """Exchange CUDA IPC handles of both state buffers so that a pass can store into any rank."""
        import ctypes

        dist = self._dist
        lib, h = self.eng._lib, self.eng._h
        mine = []
        for which in (0, 1):
            buf = ctypes.create_string_buffer(64)
            self.eng._check(lib.xyk_ipc_export(h, which, buf))
            mine.append(buf.raw)
        everyone = [None] * self.world_size
        dist.all_gather_object(everyone, mine, group=self.group)
        for r in range(self.world_size):
            for which in (0, 1):
                self.eng._check(lib.xyk_ipc_import(h, r, which, everyone[r][which]))
        dist.barrier(group=self.group)

    def close(self) -> None:
        self.eng.close()

    # -- plumbing ---------------------------------------------------------------------------------
    def _exchange(self) -> None:
        """The all-to-all the engine asked for: block s of the current buffer goes to rank s."""
        torch, dist = self._torch, self._dist
        lib, h = self.eng._lib, self.eng._h
        src, dst, bpb = c_void_p(), c_void_p(), c_size_t()
        self.eng._check(lib.xyk_exchange_info(h, byref(src), byref(dst), byref(bpb)))
        n_doubles = bpb.value * self.world_size // 8
        tin = torch.as_tensor(_DevicePtr(src.value, n_doubles), device=f"cuda:{self.device}")
        tout = torch.as_tensor(_DevicePtr(dst.value, n_doubles), device=f"cuda:{self.device}")
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        dist.all_to_all_single(tout, tin, group=self.group)
        ev1.record()
        torch.cuda.synchronize(self.device)
        self.exchange_seconds += ev0.elapsed_time(ev1) * 1e-3
        self.exchange_log.append(ev0.elapsed_time(ev1))
        self.exchange_bytes += bpb.value * (self.world_size - 1)
        self.exchanges += 1
        self.eng._check(lib.xyk_exchange_done(h))

    def _drive(self, rc: int) -> None:
        lib, h = self.eng._lib, self.eng._h
        while rc in (_lib.XYK_NEED_EXCHANGE, _lib.XYK_NEED_BARRIER):
            if rc == _lib.XYK_NEED_EXCHANGE:
                self._exchange()
            else:  # fence around a pass that stores into the peers' buffers (the engine has synced its stream)
                self._dist.barrier(group=self.group)
                self.barriers += 1
            rc = self.eng._check(lib.xyk_continue(h))

    # -- problem / evolution ------------------------------------------------------------------------
    def set_problem(self, K, omega) -> None:
        self.eng.set_problem(K, omega)

    def init_product_state(self) -> None:
        self.eng.init_product_state()

    def apply_layers(self, delta: float, order: int = 1, reps: int = 1) -> None:
        self._drive(self.eng.apply_layers(delta, order, reps))

    def global_qubits(self) -> list[int]:
        lay = self.eng.layout()
        return [q for q in range(self.n) if lay[q] >= self.n_local]

    def bring_local(self, qubits: list[int]) -> None:
        """Make every qubit of ``qubits`` local (at most one exchange)."""
        glob = self.global_qubits()
        if not any(q in glob for q in qubits):
            return
        keep = set(qubits) | set(glob)
        outgoing = [q for q in range(self.n) if q not in keep][: self.g]
        if len(outgoing) < self.g:
            raise ValueError("not enough local qubits to swap out")
        arr = (c_int * self.g)(*sorted(outgoing))
        self._drive(self.eng._check(self.eng._lib.xyk_begin_exchange(self.eng._h, arr)))

    # -- observables --------------------------------------------------------------------------------
    def _allreduce(self, vec: np.ndarray) -> np.ndarray:
        torch, dist = self._torch, self._dist
        t = torch.as_tensor(vec, dtype=torch.float64, device=f"cuda:{self.device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def xy_expectations(self) -> tuple[np.ndarray, np.ndarray]:
        """<X_q>, <Y_q> for every qubit (pauli.rs:180-214 semantics) of the full sharded state."""
        glob = self.global_qubits()
        ex, ey = self.eng.xy_expectations()  # partial sums over this shard; 0 for sharded qubits
        if glob and self._peers_mapped:
            # sharded qubits: every rank with the qubit's rank bit clear reads its partner's shard through the peer mapping
            # (NVLink) -- no qubit exchange, the sharding stays as it is; the fence in front makes every shard final, the
            # all-reduce behind it doubles as the fence that keeps the shards untouched until every reader is done
            self.eng.synchronize()
            self._dist.barrier(group=self.group)
            gx, gy = self.eng.xy_expectations_peer()
            tot = self._allreduce(np.concatenate([ex + gx, ey + gy]))
            return tot[: self.n].copy(), tot[self.n:].copy()
        tot = self._allreduce(np.concatenate([ex, ey]))
        ex, ey = tot[: self.n].copy(), tot[self.n:].copy()
        if glob:
            lay = self.eng.layout()
            entry = sorted(glob, key=lambda q: lay[q])  # rank bit k carried entry[k]
            self.bring_local(glob)
            ex2, ey2 = self.eng.xy_expectations()
            tot2 = self._allreduce(np.concatenate([ex2, ey2]))
            for q in glob:
                ex[q] = tot2[q]
                ey[q] = tot2[self.n + q]
            # back to the sharding found on entry: the evolution reuses the program compiled for it
            arr = (c_int * self.g)(*entry)
            self._drive(self.eng._check(self.eng._lib.xyk_begin_exchange(self.eng._h, arr)))
        return ex, ey

    def order_parameter(self) -> tuple[float, float]:
        ex, ey = self.xy_expectations()
        z = complex(ex.sum(), ey.sum()) / self.n
        return float(abs(z)), float(np.angle(z))

    def z_expectations(self) -> np.ndarray:
        return self._allreduce(self.eng.z_expectations())

    def norm2(self) -> float:
        return float(self._allreduce(np.array([self.eng.norm2()]))[0])

    # -- state access (tests, small n) ---------------------------------------------------------------
    def canonicalise(self) -> None:
        """Sharded qubits = the g highest logical qubits in order; local qubits ascending."""
        want = list(range(self.n - self.g, self.n))
        if self.global_qubits() != want or [self.eng.layout()[q] - self.n_local for q in want] != list(range(self.g)):
            self.bring_local(want)  # now all of `want` is local
            arr = (c_int * self.g)(*want)
            self._drive(self.eng._check(self.eng._lib.xyk_begin_exchange(self.eng._h, arr)))
        self.eng.canonicalise()

    def gather_state(self) -> np.ndarray | None:
        """Full state in canonical order on rank 0 (None elsewhere).  For tests at small n only."""
        torch, dist = self._torch, self._dist
        self.canonicalise()
        shard = self.eng.get_state()
        t = torch.as_tensor(shard.view(np.float64), device=f"cuda:{self.device}")
        parts = [torch.empty_like(t) for _ in range(self.world_size)] if self.rank == 0 else None
        dist.gather(t, parts, dst=0, group=self.group)
        if self.rank != 0:
            return None
        return np.concatenate([p.cpu().numpy() for p in parts]).view(np.complex128)

    def stats(self) -> dict:
        st = self.eng.stats()
        st.update(exchanges=self.exchanges, exchange_seconds=self.exchange_seconds, exchange_bytes=self.exchange_bytes,
                  barriers=self.barriers, fused_exchanges=self.barriers // 2)
        return st
