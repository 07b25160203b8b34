"""ctypes wrapper over ``oracle/libfrm_oracle.so`` (the C restatement of Path P).

TEST INFRASTRUCTURE — imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of ``bench.py``.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfrm_oracle.so")

STATUS = {0: "ok", 1: "no_events", 2: "ZeroDivisionError", 3: "ValueError", 4: "replay_exhausted"}


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "frm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libfrm_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        p, i32, i64, u32, u64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
        L.hfo_lattice_alloc.restype = p
        L.hfo_lattice_alloc.argtypes = [i32, i32, i32]
        L.hfo_lattice_free.argtypes = [p]
        L.hfo_lattice_set.argtypes = [p, p, p, p, p, p]
        L.hfo_lattice_get.argtypes = [p, p, p, p, p, p]
        L.hfo_lattice_species_draws.restype = i64
        L.hfo_lattice_species_draws.argtypes = [p]
        L.hfo_lattice_generate.restype = i32
        L.hfo_lattice_generate.argtypes = [p, p, u64, u32]
        L.hfo_traj_create.restype = p
        L.hfo_traj_create.argtypes = [p, i32, f64, i32, u64, u32, p, u64, p]
        L.hfo_traj_free.argtypes = [p]
        L.hfo_traj_run.restype = i64
        L.hfo_traj_run.argtypes = [p, i64, p, p, p, p, p, p, p, p]
        L.hfo_traj_counts.argtypes = [p, p]
        L.hfo_traj_tallies.argtypes = [p, p, p]
        L.hfo_traj_positions.argtypes = [p, i32, p]
        L.hfo_traj_events.argtypes = [p, i32, p, p, p, p, p]
        L.hfo_traj_flags.argtypes = [p, p]
        L.hfo_traj_cache.restype = i32
        L.hfo_traj_cache.argtypes = [p, p, p, p, p, p]
        L.hfo_traj_time_move.restype = f64
        L.hfo_traj_time_move.argtypes = [p, C.c_int32, C.c_int32, i32, p]
        L.hfo_neighbourhood.restype = i32
        L.hfo_neighbourhood.argtypes = [p, i32, C.c_int32, p, i32]
        L.hfo_run_batch.argtypes = [p, i32, f64, i32, u64, i64, i64, i64, i32, p]
        L.hfo_traj_set_excitons.argtypes = [p, i32, i32] + [p] * 11
        L.hfo_traj_excitons.restype = i32
        L.hfo_traj_excitons.argtypes = [p] * 8
        L.hfo_traj_channels.argtypes = [p, p, p]
        L.hfo_run_batch_x.argtypes = [p, i32, f64, i32, u64, i64, i64, i64, i32, i32, i32] + [p] * 12
        L.hfo_contract_uniform.restype = f64
        L.hfo_contract_uniform.argtypes = [u64, u32, u32, u64]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def resolve_proportions(dims, props):
    """Quirk Q1 (reseau.py:143-145), evaluated with Python's own ``sum`` like the reference."""
    props = tuple(float(p) for p in props)      # builtin floats: sum() is compensated in 3.12
    dims = tuple(int(d) for d in dims)
    if sum(props) != 1.:
        norm = sum(dims)
        props = (dims[0] / norm, dims[1] / norm, dims[2] / norm)
    return tuple(float(p) for p in props)


class OracleLattice:
    def __init__(self, dims):
        self.dims = tuple(int(d) for d in dims)
        self.N = self.dims[0] * self.dims[1] * self.dims[2]
        self._h = lib().hfo_lattice_alloc(*self.dims)

    @classmethod
    def generate(cls, dims, props, seed, realisation=0, resolved=False):
        """``resolved=True``: props are taken as they are (already past quirk Q1)."""
        self = cls(dims)
        pr = np.array(props if resolved else resolve_proportions(dims, props), dtype=np.float64)
        st = lib().hfo_lattice_generate(self._h, _ptr(pr), int(seed), int(realisation))
        if st != 0:
            raise ValueError(f"lattice generation failed: {STATUS[st]}")
        return self

    @classmethod
    def from_arrays(cls, dims, species, homo, lumo, s1=None, t1=None):
        self = cls(dims)
        species = np.ascontiguousarray(species, dtype=np.uint8)
        homo = np.ascontiguousarray(homo, dtype=np.float64)
        lumo = np.ascontiguousarray(lumo, dtype=np.float64)
        s1 = None if s1 is None else np.ascontiguousarray(s1, dtype=np.float64)
        t1 = None if t1 is None else np.ascontiguousarray(t1, dtype=np.float64)
        assert species.size == homo.size == lumo.size == self.N
        lib().hfo_lattice_set(self._h, _ptr(species), _ptr(homo), _ptr(lumo),
                              None if s1 is None else _ptr(s1), None if t1 is None else _ptr(t1))
        return self

    def arrays(self):
        out = dict(species=np.empty(self.N, np.uint8), homo=np.empty(self.N), lumo=np.empty(self.N),
                   s1=np.empty(self.N), t1=np.empty(self.N))
        lib().hfo_lattice_get(self._h, *(_ptr(out[k]) for k in ("species", "homo", "lumo", "s1", "t1")))
        return out

    @property
    def species_draws(self):
        return int(lib().hfo_lattice_species_draws(self._h))

    def neighbourhood(self, site, distance=1):
        buf = np.empty(4096, np.int32)
        n = lib().hfo_neighbourhood(self._h, distance, int(site), _ptr(buf), buf.size)
        if n < 0:
            raise IndexError("neighbour outside the grid")
        return buf[:n].copy()

    def run_batch_x(self, charges, field, seed, first_trajectory, n_trajectories, stop, model, n_threads=1, distance=1):
        """run_batch with integrated excitons (own model): ``model`` is an ``ExcitonModel``."""
        out = np.zeros(15, np.int64)
        lib().hfo_run_batch_x(self._h, charges, float(field), distance, int(seed), int(first_trajectory),
                              int(n_trajectories), int(stop), int(n_threads), *model.args(), _ptr(out))
        names = ("emission", "recombination", "injection", "events", "stopped_early", "forster", "dexter", "isc", "risc",
                 "recomb_host", "recomb_tadf", "recomb_fluo", "emit_host", "emit_tadf", "emit_fluo")
        return {k: int(v) for k, v in zip(names, out)}

    def run_batch(self, charges, field, seed, first_trajectory, n_trajectories, stop, n_threads=1, distance=1):
        out = np.zeros(5, np.int64)
        lib().hfo_run_batch(self._h, charges, float(field), distance, int(seed), int(first_trajectory),
                            int(n_trajectories), int(stop), int(n_threads), _ptr(out))
        return dict(emission=int(out[0]), recombination=int(out[1]), injection=int(out[2]),
                    events=int(out[3]), stopped_early=int(out[4]))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hfo_lattice_free(self._h)
            self._h = None


class ExcitonModel:
    """Integrated exciton dynamics (own model, frm_oracle.c header).  mode 1: decay at once through the exciton
    slots (the reference's behaviour); mode 2: dynamics with the exciton path's parameters, tables and tagged
    Boltzmann weights (the arrays of oracle/exciton_oracle.py or the device's own)."""

    def __init__(self, mode, params=None, tables=None, ws1=None, wt1=None):
        self.mode = int(mode)
        if self.mode == 2:
            f = lambda a, t=np.float64: np.ascontiguousarray(a, t)
            self.off = f(tables["offsets"], np.int8)
            self.g6, self.gd, self.kf, self.kd = f(tables["g6"]), f(tables["gd"]), f(tables["kf"]), f(tables["kd"])
            self.k_isc, self.k_risc = f(params["k_isc"]), f(params["k_risc"])
            self.k_rad, self.k_nonrad = f(params["k_rad"]), f(params["k_nonrad"])
            self.ws1, self.wt1 = f(ws1), f(wt1)

    def args(self):
        if self.mode != 2:
            return [self.mode, 0] + [None] * 11
        return [self.mode, int(self.g6.size)] + [_ptr(a) for a in (self.off, self.g6, self.gd, self.kf, self.kd, self.k_isc,
                                                                  self.k_risc, self.k_rad, self.k_nonrad, self.ws1, self.wt1)]


class OracleTrajectory:
    """One reference ``Lattice`` instance's dynamic state on a given realisation."""

    def __init__(self, lattice: OracleLattice, charges, seed, trajectory=0, field=0.1, distance=1, replay=None):
        self.lattice = lattice
        self._replay = None if replay is None else np.ascontiguousarray(replay, dtype=np.float64)
        st = C.c_int(0)
        self._h = lib().hfo_traj_create(lattice._h, int(charges), float(field), int(distance), int(seed),
                                        int(trajectory), None if self._replay is None else _ptr(self._replay),
                                        0 if self._replay is None else self._replay.size, C.byref(st))
        self.status = STATUS[st.value]
        self._model = None

    def set_excitons(self, model: "ExcitonModel"):
        """Integrated excitons (own model); call before the first run."""
        self._model = model                                   # keeps the arrays alive
        lib().hfo_traj_set_excitons(self._h, *model.args())

    def excitons(self):
        n = lib().hfo_traj_excitons(self._h, None, None, None, None, None, None, None)
        a = {k: np.zeros(n, np.int32) for k in ("site", "spin", "kind", "final", "rad", "onflags")}
        a["tau"] = np.zeros(n, np.float64)
        if n:
            lib().hfo_traj_excitons(self._h, *(_ptr(a[k]) for k in ("site", "spin", "kind", "final", "tau", "rad", "onflags")))
        return a

    def channels(self):
        out = np.zeros(12, np.int64)
        clock = C.c_double(0)
        lib().hfo_traj_channels(self._h, _ptr(out), C.byref(clock))
        names = ("forster", "dexter", "isc", "risc", "recomb_host", "recomb_tadf", "recomb_fluo",
                 "emit_host", "emit_tadf", "emit_fluo", "xdraws")
        d = {k: int(v) for k, v in zip(names, out)}
        d["clock"] = clock.value
        return d

    def run(self, stop, trace=True):
        stop = int(stop)
        st = C.c_int(0)
        if trace:
            tr = dict(kind=np.zeros(stop, np.uint8), particule=np.zeros(stop, np.uint8),
                      initial=np.zeros(stop, np.int32), final=np.zeros(stop, np.int32),
                      tau=np.zeros(stop, np.float64), draws=np.zeros(stop, np.int64),
                      spins=np.zeros(stop, np.int64))
            n = lib().hfo_traj_run(self._h, stop, *(_ptr(tr[k]) for k in
                                                     ("kind", "particule", "initial", "final", "tau", "draws", "spins")),
                                   C.byref(st))
            tr = {k: v[:n] for k, v in tr.items()}
        else:
            n = lib().hfo_traj_run(self._h, stop, None, None, None, None, None, None, None, C.byref(st))
            tr = None
        self.status = STATUS[st.value]
        return int(n), tr

    def counts(self):
        out = np.zeros(8, np.int64)
        lib().hfo_traj_counts(self._h, _ptr(out))
        return dict(zip(("electrons", "holes", "excitons", "move_e", "move_h", "decay", "binding", "capture"),
                        (int(v) for v in out)))

    def tallies(self):
        out = np.zeros(6, np.int64)
        iqe = C.c_double(0)
        lib().hfo_traj_tallies(self._h, _ptr(out), C.byref(iqe))
        d = dict(zip(("emission", "recombination", "injection", "step", "draws", "spins"), (int(v) for v in out)))
        d["IQE"] = iqe.value
        return d

    def positions(self, which):
        idx = {"electrons": 0, "holes": 1, "excitons": 2}[which]
        out = np.zeros(self.counts()[which], np.int32)
        lib().hfo_traj_positions(self._h, idx, _ptr(out))
        return out

    def events(self, which):
        idx = {"move_e": 0, "move_h": 1, "decay": 2, "binding": 3, "capture": 4}[which]
        n = self.counts()[which]
        ini, fin, tau = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        kind, part = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lib().hfo_traj_events(self._h, idx, _ptr(ini), _ptr(fin), _ptr(tau), _ptr(kind), _ptr(part))
        return dict(initial=ini, final=fin, tau=tau, kind=kind, particule=part)

    def flags(self):
        out = np.zeros(self.lattice.N, np.uint8)
        lib().hfo_traj_flags(self._h, _ptr(out))
        return out

    def cache(self):
        ini, fin, tau = np.zeros(10, np.int32), np.zeros(10, np.int32), np.zeros(10, np.float64)
        kind, part = np.zeros(10, np.int32), np.zeros(10, np.int32)
        n = lib().hfo_traj_cache(self._h, _ptr(ini), _ptr(fin), _ptr(tau), _ptr(kind), _ptr(part))
        return dict(initial=ini[:n], final=fin[:n], tau=tau[:n], kind=kind[:n], particule=part[:n])

    def time_move(self, initial_site, final_site, particule):
        zd = C.c_int(0)
        tau = lib().hfo_traj_time_move(self._h, int(initial_site), int(final_site), int(particule), C.byref(zd))
        if zd.value:
            raise ZeroDivisionError("float division by zero")
        return tau

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hfo_traj_free(self._h)
            self._h = None


def contract_uniform(seed, stream, ident, d):
    return lib().hfo_contract_uniform(int(seed), int(stream), int(ident), int(d))
