"""ctypes wrapper over ``oracle/libexciton_oracle.so`` — the CPU *specification* of the exciton
path (the reference has no exciton dynamics: parity unpinned, see exciton_oracle.c).

TEST INFRASTRUCTURE: imported only by ``tests/`` and the CPU-baseline legs of ``bench.py``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libexciton_oracle.so")
DENSE_LIB_PATH = os.path.join(HERE, "libexciton_dense_oracle.so")
XDT_NAMES = ("generated", "events", "forster", "dexter", "isc", "risc", "rad_host", "rad_tadf", "rad_fluo",
             "nonrad_s_host", "nonrad_s_tadf", "nonrad_s_fluo", "nonrad_t_host", "nonrad_t_tadf", "nonrad_t_fluo", "_15",
             "ann_ss", "ann_st", "ann_ts", "ann_tt", "tta_up", "_21", "_22", "_23")
XT_NAMES = ("generated", "events", "forster", "dexter", "isc", "risc", "rad_host", "rad_tadf", "rad_fluo",
            "nonrad_s_host", "nonrad_s_tadf", "nonrad_s_fluo", "nonrad_t_host", "nonrad_t_tadf", "nonrad_t_fluo", "_pad")
KT = 8.617333262 * 10.**(-5) * 300.

_lib = None


def build(force=False):
    src = os.path.join(HERE, "exciton_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libexciton_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


_dense = None


def dense_lib():
    global _dense
    if _dense is None:
        src = os.path.join(HERE, "exciton_dense_oracle.c")
        if not os.path.exists(DENSE_LIB_PATH) or os.path.getmtime(DENSE_LIB_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE, "-B", "libexciton_dense_oracle.so"], check=True, capture_output=True)
        _dense = C.CDLL(DENSE_LIB_PATH)
        _dense.hfxd_run.restype = C.c_int64
    return _dense


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params():
    """The defaults of hf_x_default_params (include/hfkmc.h), restated."""
    return dict(cutoff=4.0,
                forster_radius=np.array([[1.0, 3.0, 3.5], [0.5, 2.0, 4.0], [0.5, 1.0, 2.5]]),
                dexter_prefactor=np.full((3, 3), 1e10), dexter_length=0.5,
                k_isc=np.array([1e5, 1e7, 1e5]), k_risc=np.array([0.0, 1e7, 0.0]),
                k_rad=np.array([[0.0, 0.0], [1e7, 0.0], [2e8, 0.0]]),
                k_nonrad=np.array([[2e8, 1e3], [1e6, 1e4], [1e7, 1e5]]),
                singlet_fraction=0.25, outcoupling=0.2)


def tables(params):
    """Neighbour / prefactor tables from the parameters, independently of the device code."""
    R = int(math.floor(params["cutoff"]))
    off, g6, gd = [], [], []
    for dz in range(-R, R + 1):
        for dy in range(-R, R + 1):
            for dx in range(-R, R + 1):
                d2 = dx * dx + dy * dy + dz * dz
                if d2 == 0 or d2 > params["cutoff"] ** 2:
                    continue
                off.append((dx, dy, dz))
                g6.append(1.0 / ((float(d2) * float(d2)) * float(d2)))
                gd.append(math.exp(-2.0 * math.sqrt(float(d2)) / params["dexter_length"]))
    kf = np.zeros((3, 3))
    for a in range(3):
        ks = params["k_rad"][a][0] + params["k_nonrad"][a][0]
        for b in range(3):
            r0 = params["forster_radius"][a][b]
            r3 = (r0 * r0) * r0
            kf[a, b] = ks * (r3 * r3)
    return dict(offsets=np.array(off, dtype=np.int8), g6=np.array(g6), gd=np.array(gd), kf=kf,
                kd=np.array(params["dexter_prefactor"], dtype=np.float64))


def tag(w, species):
    """Replace the two least-significant mantissa bits of each weight by the species code."""
    bits = np.ascontiguousarray(w, np.float64).view(np.uint64)
    return ((bits & ~np.uint64(3)) | np.asarray(species, np.uint64)).view(np.float64)


def weights(s1, t1, species):
    """Tagged Boltzmann weights exp(-E/kT) per site with the host libm (the device computes its own
    with CUDA's exp; tests pass the device's arrays to the oracle so that both sides start from
    identical inputs)."""
    return tag(np.exp(-(np.asarray(s1) / KT)), species), tag(np.exp(-(np.asarray(t1) / KT)), species)


class ExcitonOracle:
    def __init__(self, dims, species, ws1, wt1, params, tab=None):
        """ws1 / wt1 are the TAGGED weights; species is kept only for reference."""
        self.dims = tuple(int(d) for d in dims)
        self.species = np.ascontiguousarray(species, np.uint8)
        self.ws1 = np.ascontiguousarray(ws1, np.float64)
        self.wt1 = np.ascontiguousarray(wt1, np.float64)
        self.params = params
        tab = tab or tables(params)
        self.off = np.ascontiguousarray(tab["offsets"], np.int8)
        self.g6 = np.ascontiguousarray(tab["g6"], np.float64)
        self.gd = np.ascontiguousarray(tab["gd"], np.float64)
        self.kf = np.ascontiguousarray(tab["kf"], np.float64)
        self.kd = np.ascontiguousarray(tab["kd"], np.float64)
        self.k_isc = np.ascontiguousarray(params["k_isc"], np.float64)
        self.k_risc = np.ascontiguousarray(params["k_risc"], np.float64)
        self.k_rad = np.ascontiguousarray(params["k_rad"], np.float64)
        self.k_nonrad = np.ascontiguousarray(params["k_nonrad"], np.float64)

    def _model_args(self):
        X, Y, Z = self.dims
        return [C.c_int(X), C.c_int(Y), C.c_int(Z), _p(self.species), _p(self.ws1), _p(self.wt1), C.c_int(len(self.g6)),
                _p(self.off), _p(self.g6), _p(self.gd), _p(self.kf), _p(self.kd), _p(self.k_isc), _p(self.k_risc),
                _p(self.k_rad), _p(self.k_nonrad), C.c_double(self.params["singlet_fraction"])]

    def run(self, seed, trajectory, n_events):
        n = int(n_events)
        tr = dict(kind=np.zeros(n, np.uint8), initial=np.zeros(n, np.int32), final=np.zeros(n, np.int32),
                  dt=np.zeros(n), channel=np.zeros(n, np.int32), spin=np.zeros(n, np.uint8))
        tallies = np.zeros(16, np.uint64)
        state = np.zeros(3, np.int64)
        life = C.c_double(0)
        lib().hfx_run(*self._model_args(), C.c_uint64(seed), C.c_uint32(trajectory), C.c_int64(n),
                      _p(tr["kind"]), _p(tr["initial"]), _p(tr["final"]), _p(tr["dt"]), _p(tr["channel"]), _p(tr["spin"]),
                      _p(tallies), _p(state), C.byref(life))
        return dict(trace=tr, tallies={k: int(v) for k, v in zip(XT_NAMES, tallies) if not k.startswith("_")},
                    site=int(state[0]), spin=int(state[1]), draws=int(state[2]), lifetime_sum=life.value)

    def run_batch(self, seed, first, n, n_events, n_threads=1):
        tallies = np.zeros(16, np.uint64)
        lib().hfx_run_batch(*self._model_args(), C.c_uint64(seed), C.c_int64(first), C.c_int64(n), C.c_int64(n_events),
                            C.c_int(n_threads), _p(tallies))
        return {k: int(v) for k, v in zip(XT_NAMES, tallies) if not k.startswith("_")}


class DenseExcitonOracle(ExcitonOracle):
    """The high-density model (oracle/exciton_dense_oracle.c): n interacting excitons per device."""

    def __init__(self, dims, species, ws1, wt1, params, tab=None, annihilation=((1.0, 1.0), (1.0, 1.0)), p_tta=0.25):
        super().__init__(dims, species, ws1, wt1, params, tab)
        self.ann = np.ascontiguousarray(annihilation, np.float64).reshape(4)
        self.p_tta = float(p_tta)

    def _dense_args(self):
        X, Y, Z = self.dims
        return [C.c_int(X), C.c_int(Y), C.c_int(Z), _p(self.ws1), _p(self.wt1), C.c_int(len(self.g6)), _p(self.off),
                _p(self.g6), _p(self.gd), _p(self.kf), _p(self.kd), _p(self.k_isc), _p(self.k_risc), _p(self.k_rad),
                _p(self.k_nonrad), C.c_double(self.params["singlet_fraction"]), C.c_double(self.params["cutoff"]),
                _p(self.ann), C.c_double(self.p_tta)]

    def run_device(self, seed, device, n_excitons, n_events, verify=False):
        n, m = int(n_events), int(n_excitons)
        tr = dict(kind=np.zeros(n, np.uint8), exciton=np.zeros(n, np.int32), initial=np.zeros(n, np.int32),
                  final=np.zeros(n, np.int32), channel=np.zeros(n, np.int32), dt=np.zeros(n), spin=np.zeros(n, np.uint8))
        tallies = np.zeros(24, np.uint64)
        pos, spin, tot, scal = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m), np.zeros(3)
        bad = dense_lib().hfxd_run(*self._dense_args(), C.c_uint64(seed), C.c_uint32(device), C.c_int(m), C.c_int64(n),
                                   C.c_int(int(verify)), _p(tr["kind"]), _p(tr["exciton"]), _p(tr["initial"]), _p(tr["final"]),
                                   _p(tr["channel"]), _p(tr["dt"]), _p(tr["spin"]), _p(tallies), _p(pos), _p(spin), _p(tot), _p(scal))
        return dict(trace=tr, tallies={k: int(v) for k, v in zip(XDT_NAMES, tallies) if not k.startswith("_")},
                    site=pos, spin=spin, total_rate=tot, time=scal[0], lifetime_sum=scal[1], draws=int(scal[2]),
                    cache_mismatches=int(bad))

    def run_devices(self, seed, first, n_devices, n_excitons, n_events, n_threads=1):
        tallies = np.zeros(24, np.uint64)
        dense_lib().hfxd_run_batch(*self._dense_args(), C.c_uint64(seed), C.c_int64(first), C.c_int64(n_devices),
                                   C.c_int(int(n_excitons)), C.c_int64(n_events), C.c_int(n_threads), _p(tallies))
        return {k: int(v) for k, v in zip(XDT_NAMES, tallies) if not k.startswith("_")}
