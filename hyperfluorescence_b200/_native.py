"""ctypes binding of ``libhfkmc.so`` (C-ABI in ``include/hfkmc.h``).

There is no CPU implementation behind this module: if the shared library is missing, or no
sm_100 device is visible, every entry point fails loudly.  The parity oracle under ``oracle/``
is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HFKMC_LIB") or os.path.join(HERE, "libhfkmc.so")   # HFKMC_LIB: kernel-variant experiments

HF_N_TALLIES = 24
TALLY_NAMES = ("emission", "recombination", "injection", "steps", "fired", "move_e", "move_h", "bound", "decay",
               "capture_e", "capture_h", "recomb_host", "recomb_tadf", "recomb_fluo", "emit_tadf", "emit_fluo", "stopped",
               "x_forster", "x_dexter", "x_isc", "x_risc")

HF_FLAG_NONE, HF_FLAG_WARP_KERNEL = 0, 1     # hf_config.flags: force the warp-per-trajectory kernel (cross-checks)
HF_OK, HF_ERR_BAD_ARG, HF_ERR_CUDA, HF_ERR_NO_DEVICE, HF_ERR_UNSUPPORTED, HF_ERR_STATE = range(6)
TRAJ_OK, TRAJ_NO_EVENTS, TRAJ_ZERO_DIVISION, TRAJ_VALUE_ERROR, TRAJ_REPLAY_EXHAUSTED, TRAJ_CAPACITY = range(6)


class HfError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libhfkmc error {code}: {message}")
        self.code = code


class HfConfig(C.Structure):
    _fields_ = [("dims", C.c_int32 * 3), ("props", C.c_double * 3), ("electric_field", C.c_double),
                ("charges", C.c_int32), ("distance", C.c_int32), ("n_trajectories", C.c_int64),
                ("n_realisations", C.c_int32), ("first_trajectory", C.c_uint32), ("first_realisation", C.c_uint32),
                ("seed", C.c_uint64), ("flags", C.c_uint32), ("device", C.c_int32)]


class HfEvent(C.Structure):
    _fields_ = [("initial", C.c_int32), ("final", C.c_int32), ("tau", C.c_double), ("kind", C.c_uint8),
                ("particule", C.c_uint8), ("pad", C.c_uint8 * 2), ("spins", C.c_uint32), ("draws", C.c_uint64)]


EVENT_DTYPE = np.dtype([("initial", "<i4"), ("final", "<i4"), ("tau", "<f8"), ("kind", "u1"), ("particule", "u1"),
                        ("pad", "u1", (2,)), ("spins", "<u4"), ("draws", "<u8")])
assert EVENT_DTYPE.itemsize == C.sizeof(HfEvent) == 32


class HfTrajInfo(C.Structure):
    _fields_ = [("n_electrons", C.c_int32), ("n_holes", C.c_int32), ("n_eflags", C.c_int32), ("n_hflags", C.c_int32),
                ("n_move_e", C.c_int32), ("n_move_h", C.c_int32), ("special_kind", C.c_int32),
                ("special_particule", C.c_int32), ("special_site", C.c_int32), ("exciton_state", C.c_int32),
                ("status", C.c_int32), ("cache_count", C.c_int32), ("draws", C.c_uint64), ("spins", C.c_uint64),
                ("emission", C.c_uint64), ("recombination", C.c_uint64), ("injection", C.c_uint64),
                ("steps", C.c_uint64), ("fired", C.c_uint64)]


class HfXParams(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("forster_radius", (C.c_double * 3) * 3),
                ("dexter_prefactor", (C.c_double * 3) * 3), ("dexter_length", C.c_double),
                ("k_isc", C.c_double * 3), ("k_risc", C.c_double * 3), ("k_rad", (C.c_double * 2) * 3),
                ("k_nonrad", (C.c_double * 2) * 3), ("singlet_fraction", C.c_double), ("outcoupling", C.c_double)]


HF_XT_N = 16
XT_NAMES = ("generated", "events", "forster", "dexter", "isc", "risc", "rad_host", "rad_tadf", "rad_fluo",
            "nonrad_s_host", "nonrad_s_tadf", "nonrad_s_fluo", "nonrad_t_host", "nonrad_t_tadf", "nonrad_t_fluo")

HF_XDT_N = 24
XDT_NAMES = XT_NAMES + ("_15", "ann_ss", "ann_st", "ann_ts", "ann_tt", "tta_up", "_21", "_22", "_23")

INFO_DTYPE = np.dtype([(n, "<i4") for n in ("n_electrons", "n_holes", "n_eflags", "n_hflags", "n_move_e", "n_move_h",
                                             "special_kind", "special_particule", "special_site", "exciton_state",
                                             "status", "cache_count")] +
                      [(n, "<u8") for n in ("draws", "spins", "emission", "recombination", "injection", "steps", "fired")])
assert INFO_DTYPE.itemsize == C.sizeof(HfTrajInfo)

# name -> (restype, argtypes): every symbol include/hfkmc.h declares
P, I32, I64, F64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SIGNATURES = {
    "hf_last_error": (C.c_char_p, []),
    "hf_abi_version": (C.c_int, []),
    "hf_device_count": (C.c_int, []),
    "hf_create": (C.c_int, [C.POINTER(HfConfig), C.POINTER(P)]),
    "hf_destroy": (C.c_int, [P]),
    "hf_set_stream": (C.c_int, [P, P]),
    "hf_sync": (C.c_int, [P]),
    "hf_generate_lattice": (C.c_int, [P]),
    "hf_generate_lattice_mixed": (C.c_int, [P, P]),
    "hf_read_realisation_tallies": (C.c_int, [P, P]),
    "hf_read_infos": (C.c_int, [P, I64, I64, P]),
    "hf_upload_lattice": (C.c_int, [P, I32, P, P, P, P, P]),
    "hf_download_lattice": (C.c_int, [P, I32, P, P, P, P, P]),
    "hf_inject": (C.c_int, [P]),
    "hf_set_replay": (C.c_int, [P, P, I64]),
    "hf_run": (C.c_int, [P, I64]),
    "hf_read_tallies": (C.c_int, [P, P]),
    "hf_reduce_tallies_device": (C.c_int, [P, P]),
    "hf_read_info": (C.c_int, [P, I64, C.POINTER(HfTrajInfo)]),
    "hf_read_positions": (C.c_int, [P, I64, I32, P, C.POINTER(I32)]),
    "hf_read_flags": (C.c_int, [P, I64, I32, P, C.POINTER(I32)]),
    "hf_p_excitons": (C.c_int, [P, I32]),
    "hf_read_excitons": (C.c_int, [P, I64, P, P, P, P, P, P, C.POINTER(I32), C.POINTER(C.c_uint64), C.POINTER(F64)]),
    "hf_read_move_events": (C.c_int, [P, I64, I32, P, P, P, C.POINTER(I32)]),
    "hf_read_last_events": (C.c_int, [P, I64, P, C.POINTER(I32)]),
    "hf_set_trace": (C.c_int, [P, I64, I64, I64]),
    "hf_read_trace": (C.c_int, [P, I64, P, C.POINTER(I64)]),
    "hf_eval_hops": (C.c_int, [P, I64, I32, I32, P, P, P, P, P, P]),
    "hf_run_timing": (C.c_int, [P, C.POINTER(F64), C.POINTER(I64)]),
    "hf_x_default_params": (C.c_int, [C.POINTER(HfXParams)]),
    "hf_x_configure": (C.c_int, [P, C.POINTER(HfXParams)]),
    "hf_x_spawn": (C.c_int, [P]),
    "hf_x_run": (C.c_int, [P, I64]),
    "hf_x_read_tallies": (C.c_int, [P, P]),
    "hf_x_read_realisation_tallies": (C.c_int, [P, P]),
    "hf_x_read_tables": (C.c_int, [P, C.POINTER(I32), P, P, P, P, P]),
    "hf_x_download_weights": (C.c_int, [P, I32, P, P]),
    "hf_x_launch_info": (C.c_int, [P, C.POINTER(I32), C.POINTER(I32), C.POINTER(I32)]),
    "hf_x_read_state": (C.c_int, [P, I64, I64, P, P, P, P, P]),
    "hf_x_read_trace": (C.c_int, [P, I64, P, C.POINTER(I64)]),
    "hf_xd_configure": (C.c_int, [P, I32, P, F64]),
    "hf_xd_spawn": (C.c_int, [P]),
    "hf_xd_run": (C.c_int, [P, I64]),
    "hf_xd_read_tallies": (C.c_int, [P, P]),
    "hf_xd_read_realisation_tallies": (C.c_int, [P, P]),
    "hf_xd_read_device": (C.c_int, [P, I64, P, P, P, P, P]),
    "hf_xd_read_occupancy": (C.c_int, [P, I64, P]),
}

_lib = None


def load():
    """Load libhfkmc.so and bind every declared symbol; raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m hyperfluorescence_b200.build` "
                "(hyperfluorescence_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(code):
    if code != HF_OK:
        raise HfError(code, load().hf_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Sim:
    """Owning wrapper of one ``hf_sim`` handle."""

    def __init__(self, dims, props, electric_field, charges, distance=1, n_trajectories=1, n_realisations=1,
                 seed=0, first_trajectory=0, first_realisation=0, device=0, flags=0):
        lib = load()
        cfg = HfConfig()
        cfg.dims[:] = [int(d) for d in dims]
        cfg.props[:] = [float(p) for p in props]
        cfg.electric_field = float(electric_field)
        cfg.charges, cfg.distance = int(charges), int(distance)
        self.exciton_mode = 0
        cfg.n_trajectories, cfg.n_realisations = int(n_trajectories), int(n_realisations)
        cfg.first_trajectory, cfg.first_realisation = int(first_trajectory), int(first_realisation)
        cfg.seed, cfg.flags, cfg.device = int(seed) & (2 ** 64 - 1), int(flags), int(device)
        self._h = P()
        self.dims = tuple(int(d) for d in dims)
        self.N = self.dims[0] * self.dims[1] * self.dims[2]
        self.charges = int(charges)
        self.n_trajectories, self.n_realisations = int(n_trajectories), int(n_realisations)
        check(lib.hf_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            load().hf_destroy(self._h)
            self._h = None

    __del__ = close

    # -- lattice ---------------------------------------------------------------------------
    def generate_lattice(self):
        check(load().hf_generate_lattice(self._h))

    def generate_lattice_mixed(self, props):
        """One (host, tadf, fluo) composition per realisation; props has shape (n_realisations, 3)."""
        props = np.ascontiguousarray(props, np.float64)
        if props.shape != (self.n_realisations, 3):
            raise ValueError(f"props must have shape ({self.n_realisations}, 3)")
        check(load().hf_generate_lattice_mixed(self._h, ptr(props)))

    def realisation_tallies(self):
        out = np.zeros((self.n_realisations, 5), np.uint64)
        check(load().hf_read_realisation_tallies(self._h, ptr(out)))
        return {k: out[:, i].astype(np.int64) for i, k in enumerate(("emission", "recombination", "injection", "steps", "fired"))}

    def infos(self, first=0, n=None):
        """Structured array of the scalar state of trajectories [first, first + n)."""
        n = self.n_trajectories - first if n is None else n
        out = (HfTrajInfo * n)()
        check(load().hf_read_infos(self._h, int(first), int(n), out))
        return np.frombuffer(out, dtype=INFO_DTYPE).copy()

    def upload_lattice(self, realisation, species, homo, lumo, s1=None, t1=None):
        arrs = [np.ascontiguousarray(species, np.uint8)] + [
            None if a is None else np.ascontiguousarray(a, np.float64) for a in (homo, lumo, s1, t1)]
        for a in arrs:
            if a is not None and a.size != self.N:
                raise ValueError(f"lattice array has {a.size} entries, expected {self.N}")
        check(load().hf_upload_lattice(self._h, int(realisation), *(ptr(a) for a in arrs)))

    def download_lattice(self, realisation=0):
        out = dict(species=np.empty(self.N, np.uint8), homo=np.empty(self.N), lumo=np.empty(self.N),
                   s1=np.empty(self.N), t1=np.empty(self.N))
        check(load().hf_download_lattice(self._h, int(realisation),
                                         *(ptr(out[k]) for k in ("species", "homo", "lumo", "s1", "t1"))))
        return out

    # -- dynamics ----------------------------------------------------------------------------
    def set_replay(self, u):
        if u is None:
            check(load().hf_set_replay(self._h, None, 0))
            return
        u = np.ascontiguousarray(u, np.float64)
        if u.ndim == 1:
            u = np.tile(u, (self.n_trajectories, 1)) if self.n_trajectories > 1 else u[None, :]
        if u.shape[0] != self.n_trajectories:
            raise ValueError("replay must have one row per trajectory")
        u = np.ascontiguousarray(u)
        check(load().hf_set_replay(self._h, ptr(u), u.shape[1]))

    def inject(self):
        check(load().hf_inject(self._h))

    def run(self, stop):
        check(load().hf_run(self._h, int(stop)))

    def sync(self):
        check(load().hf_sync(self._h))

    def set_stream(self, cuda_stream):
        check(load().hf_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def run_timing(self):
        ms, n = F64(0), I64(0)
        check(load().hf_run_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- read-back -----------------------------------------------------------------------------
    def tallies(self):
        out = np.zeros(HF_N_TALLIES, np.uint64)
        check(load().hf_read_tallies(self._h, ptr(out)))
        return {k: int(out[i]) for i, k in enumerate(TALLY_NAMES)}

    def reduce_tallies_device(self, device_ptr):
        check(load().hf_reduce_tallies_device(self._h, C.c_void_p(int(device_ptr))))

    def info(self, trajectory=0):
        info = HfTrajInfo()
        check(load().hf_read_info(self._h, int(trajectory), C.byref(info)))
        return info

    def _list(self, fn, trajectory, which, cap):
        out = np.zeros(max(cap, 1), np.int32)
        n = I32(out.size)
        check(fn(self._h, int(trajectory), int(which), ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def positions(self, trajectory, which):
        return self._list(load().hf_read_positions, trajectory, which, self.charges)

    def flags(self, trajectory, which):
        return self._list(load().hf_read_flags, trajectory, which, 2 * self.charges + 8)

    def move_events(self, trajectory, which):
        cap = max(self.charges, 1)
        ini, fin, tau = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
        n = I32(cap)
        check(load().hf_read_move_events(self._h, int(trajectory), int(which), ptr(ini), ptr(fin), ptr(tau), C.byref(n)))
        return ini[:n.value].copy(), fin[:n.value].copy(), tau[:n.value].copy()

    def last_events(self, trajectory=0):
        out = np.zeros(10, EVENT_DTYPE)
        n = I32(10)
        check(load().hf_read_last_events(self._h, int(trajectory), ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def set_trace(self, first, n, capacity):
        check(load().hf_set_trace(self._h, int(first), int(n), int(capacity)))

    def trace(self, trajectory, capacity):
        out = np.zeros(max(int(capacity), 1), EVENT_DTYPE)
        n = I64(out.size)
        check(load().hf_read_trace(self._h, int(trajectory), ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def eval_hops(self, trajectory, particule, initial, final, u):
        initial = np.ascontiguousarray(initial, np.int32)
        final = np.ascontiguousarray(final, np.int32)
        u = np.ascontiguousarray(u, np.float64)
        n = initial.size
        tau, rate, zd = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
        check(load().hf_eval_hops(self._h, int(trajectory), int(particule), n, ptr(initial), ptr(final), ptr(u),
                                  ptr(tau), ptr(rate), ptr(zd)))
        return tau, rate, zd

    # -- integrated excitons (own model in the reference's slots) ---------------------------------
    def p_excitons(self, mode):
        """0: off (the reference as it is); 1: decay at once through the exciton slots; 2: dynamics with the
        parameters of ``x_configure`` (call it first).  Needs a new ``inject``."""
        check(load().hf_p_excitons(self._h, int(mode)))
        self.exciton_mode = int(mode)

    def excitons(self, trajectory=0):
        cap = max(self.charges, 1)
        a = {k: np.zeros(cap, np.int32) for k in ("site", "spin", "kind", "final", "radiative")}
        tau = np.zeros(cap, np.float64)
        n, draws, clock = I32(cap), C.c_uint64(0), F64(0.0)
        check(load().hf_read_excitons(self._h, int(trajectory), ptr(a["site"]), ptr(a["spin"]), ptr(a["kind"]), ptr(a["final"]),
                                      ptr(tau), ptr(a["radiative"]), C.byref(n), C.byref(draws), C.byref(clock)))
        out = {k: v[:n.value].copy() for k, v in a.items()}
        out.update(tau=tau[:n.value].copy(), draws=int(draws.value), clock=float(clock.value))
        return out

    # -- exciton trajectories (Path X) -----------------------------------------------------------
    def x_configure(self, params: "HfXParams | None" = None):
        if params is None:
            params = x_default_params()
        check(load().hf_x_configure(self._h, C.byref(params)))

    def x_spawn(self):
        check(load().hf_x_spawn(self._h))

    def x_run(self, n_events):
        check(load().hf_x_run(self._h, int(n_events)))

    def x_tallies(self):
        out = np.zeros(HF_XT_N, np.uint64)
        check(load().hf_x_read_tallies(self._h, ptr(out)))
        return {k: int(out[i]) for i, k in enumerate(XT_NAMES)}

    def x_realisation_tallies(self):
        out = np.zeros((self.n_realisations, HF_XT_N), np.uint64)
        check(load().hf_x_read_realisation_tallies(self._h, ptr(out)))
        return {k: out[:, i].astype(np.int64) for i, k in enumerate(XT_NAMES)}

    def x_tables(self):
        n = I32(0)
        check(load().hf_x_read_tables(self._h, C.byref(n), None, None, None, None, None))
        m = n.value
        off, g6, gd = np.zeros((m, 3), np.int8), np.zeros(m), np.zeros(m)
        kf, kd = np.zeros((3, 3)), np.zeros((3, 3))
        n = I32(m)
        check(load().hf_x_read_tables(self._h, C.byref(n), ptr(off), ptr(g6), ptr(gd), ptr(kf), ptr(kd)))
        return dict(offsets=off, g6=g6, gd=gd, kf=kf, kd=kd)

    def x_weights(self, realisation=0):
        ws1, wt1 = np.zeros(self.N), np.zeros(self.N)
        check(load().hf_x_download_weights(self._h, int(realisation), ptr(ws1), ptr(wt1)))
        return ws1, wt1

    def x_launch_info(self):
        """How hf_x_run is launched: trajectories per warp, brick-layout weights or not, grid (hf_x_launch_info)."""
        block, brick, grid = I32(0), I32(0), I32(0)
        check(load().hf_x_launch_info(self._h, C.byref(block), C.byref(brick), C.byref(grid)))
        return dict(block=block.value, brick=bool(brick.value), grid=grid.value)

    def x_state(self, first=0, n=None):
        n = self.n_trajectories - first if n is None else int(n)
        site, spin = np.zeros(n, np.int32), np.zeros(n, np.int32)
        time, life, draws = np.zeros(n), np.zeros(n), np.zeros(n, np.uint64)
        check(load().hf_x_read_state(self._h, int(first), n, ptr(site), ptr(spin), ptr(time), ptr(draws), ptr(life)))
        return dict(site=site, spin=spin, time=time, draws=draws, lifetime_sum=life)

    def x_trace(self, trajectory, capacity):
        out = np.zeros(max(int(capacity), 1), EVENT_DTYPE)
        n = I64(out.size)
        check(load().hf_x_read_trace(self._h, int(trajectory), ptr(out), C.byref(n)))
        return out[:n.value].copy()


    # -- high-density exciton devices (Path XD) ---------------------------------------------------
    def xd_configure(self, n_excitons, annihilation=((1.0, 1.0), (1.0, 1.0)), p_tta=0.25):
        ann = np.ascontiguousarray(annihilation, np.float64).reshape(4)
        self.xd_n = int(n_excitons)
        check(load().hf_xd_configure(self._h, self.xd_n, ptr(ann), float(p_tta)))

    def xd_spawn(self):
        check(load().hf_xd_spawn(self._h))

    def xd_run(self, n_events):
        check(load().hf_xd_run(self._h, int(n_events)))

    def xd_tallies(self):
        out = np.zeros(HF_XDT_N, np.uint64)
        check(load().hf_xd_read_tallies(self._h, ptr(out)))
        return {k: int(out[i]) for i, k in enumerate(XDT_NAMES) if not k.startswith("_")}

    def xd_realisation_tallies(self):
        out = np.zeros((self.n_realisations, HF_XDT_N), np.uint64)
        check(load().hf_xd_read_realisation_tallies(self._h, ptr(out)))
        return {k: out[:, i].astype(np.int64) for i, k in enumerate(XDT_NAMES) if not k.startswith("_")}

    def xd_device(self, device=0):
        n = self.xd_n
        site, spin, tot, birth, scal = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n), np.zeros(n), np.zeros(3)
        check(load().hf_xd_read_device(self._h, int(device), ptr(site), ptr(spin), ptr(tot), ptr(birth), ptr(scal)))
        return dict(site=site, spin=spin, total_rate=tot, birth=birth, time=scal[0], lifetime_sum=scal[1], draws=int(scal[2]))

    def xd_occupancy(self, device=0):
        codes = np.zeros(self.N, np.uint8)
        check(load().hf_xd_read_occupancy(self._h, int(device), ptr(codes)))
        return codes


def x_default_params() -> HfXParams:
    p = HfXParams()
    check(load().hf_x_default_params(C.byref(p)))
    return p


def x_params_to_dict(p: HfXParams) -> dict:
    return dict(cutoff=p.cutoff, forster_radius=np.array([[p.forster_radius[a][b] for b in range(3)] for a in range(3)]),
                dexter_prefactor=np.array([[p.dexter_prefactor[a][b] for b in range(3)] for a in range(3)]),
                dexter_length=p.dexter_length, k_isc=np.array(list(p.k_isc)), k_risc=np.array(list(p.k_risc)),
                k_rad=np.array([[p.k_rad[a][s] for s in range(2)] for a in range(3)]),
                k_nonrad=np.array([[p.k_nonrad[a][s] for s in range(2)] for a in range(3)]),
                singlet_fraction=p.singlet_fraction, outcoupling=p.outcoupling)


def x_params_from_dict(d: dict) -> HfXParams:
    p = x_default_params()
    for k, v in d.items():
        if k in ("cutoff", "dexter_length", "singlet_fraction", "outcoupling"):
            setattr(p, k, float(v))
        elif k in ("k_isc", "k_risc"):
            for a in range(3):
                getattr(p, k)[a] = float(v[a])
        elif k in ("forster_radius", "dexter_prefactor", "k_rad", "k_nonrad"):
            arr = np.asarray(v, dtype=np.float64)
            for a in range(arr.shape[0]):
                for b in range(arr.shape[1]):
                    getattr(p, k)[a][b] = float(arr[a, b])
        else:
            raise KeyError(k)
    return p
