"""Drive the UNMODIFIED reference under injected uniforms and record what it does.

TEST INFRASTRUCTURE — runs only where ``/root/reference`` is mounted (the build container).
It is used by ``tests/golden/make_golden.py`` to freeze golden vectors and by the optional
live cross-checks in ``tests/`` (skipped when the reference is absent, e.g. on the GPU box).
Nothing under ``hyperfluorescence_b200/`` imports it.

Method: the reference binds ``Random`` as a module global in both
``reseau`` (``reseau.py:19``) and ``molecule`` (``molecule.py:13``).  Rebinding those two names
to a factory *before* ``Lattice(...)`` is constructed makes every consumer draw from the Philox
streams of ``oracle/philox.py`` without editing a reference file.  Instance creation order is
deterministic: #0 is ``Lattice._seed`` (``reseau.py:113``), then one per molecule in z, y, x
order (``reseau.py:173``).
"""
from __future__ import annotations

import os
import random
import sys
from contextlib import contextmanager

import numpy as np

from . import philox as px

REFERENCE_DIR = os.environ.get("HF_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "reseau.py"))


def import_reference():
    """Import the reference's ``reseau`` module (and its siblings) from REFERENCE_DIR."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_DIR}")
    sys.dont_write_bytecode = True  # the tree is read-only
    for name in ("reseau", "molecule", "event", "constants"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_DIR):
            raise RuntimeError(f"module {name!r} already imported from {mod.__file__}")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import reseau  # noqa: F401  (reference module)
    import molecule  # noqa: F401
    import event  # noqa: F401
    import constants  # noqa: F401
    return sys.modules["reseau"], sys.modules["molecule"], sys.modules["event"], sys.modules["constants"]


class _Context:
    def __init__(self, seed, trajectory, realisation, replay_traj=None, replay_spin=None):
        self.seed = seed
        self.trajectory = trajectory
        self.realisation = realisation
        self.traj = px.UniformStream(seed, px.STREAM_TRAJ, trajectory, replay=replay_traj)
        self.spin = px.UniformStream(seed, px.STREAM_SPIN, trajectory, replay=replay_spin)
        self.species = px.UniformStream(seed, px.STREAM_SPECIES, realisation)
        self.instances = 0


class _LatticeRandom(random.Random):
    """Instance #0.  Species-shuffle draws come from the realisation's SPECIES stream, every
    later draw from the trajectory's TRAJ stream (the switch happens when the first molecule
    RNG is created, i.e. right after ``sample`` returned, ``reseau.py:164-173``)."""

    def __init__(self, ctx):
        self.ctx = ctx
        super().__init__(0)

    def random(self):
        ctx = self.ctx
        return (ctx.species if ctx.instances == 1 else ctx.traj).next()


class _MoleculeRandom(random.Random):
    """Instances #1.. : the first four uniforms (the four ``gauss`` calls of the species
    constructors, ``molecule.py:180-183``) come from ENERGY[4*site + j]; later draws are the
    spin draws of ``generate_exciton`` (``molecule.py:92``) and come from the trajectory's SPIN
    stream in call order."""

    def __init__(self, ctx, site):
        self.ctx, self.site, self.n = ctx, site, 0
        super().__init__(0)

    def random(self):
        ctx = self.ctx
        if self.n < 4:
            u = px.uniform(ctx.seed, px.STREAM_ENERGY, ctx.realisation, 4 * self.site + self.n)
            self.n += 1
            return u
        return ctx.spin.next()


@contextmanager
def injected(seed, trajectory=0, realisation=0, replay_traj=None, replay_spin=None):
    reseau, molecule, _, _ = import_reference()
    ctx = _Context(seed, trajectory, realisation, replay_traj, replay_spin)

    def factory(*_args):
        i = ctx.instances
        ctx.instances += 1
        return _LatticeRandom(ctx) if i == 0 else _MoleculeRandom(ctx, i - 1)

    saved = (reseau.Random, molecule.Random)
    reseau.Random = factory
    molecule.Random = factory
    try:
        yield ctx
    finally:
        reseau.Random, molecule.Random = saved


# ----------------------------------------------------------------------------------------------
# state extraction
# ----------------------------------------------------------------------------------------------

def site_index(p, dims):
    X, Y, _ = dims
    return (int(p.z) * Y + int(p.y)) * X + int(p.x)


def lattice_arrays(lat):
    """species (0 host / 1 TADF / 2 fluorescent) and the four site energies, flat z,y,x order."""
    _, molecule, _, _ = import_reference()
    code = {molecule.Host: 0, molecule.TADF: 1, molecule.Fluorescent: 2}
    mols = [m for plane in lat._grid for row in plane for m in row]
    return dict(
        species=np.array([code[type(m)] for m in mols], dtype=np.uint8),
        homo=np.array([m.homo_energy for m in mols], dtype=np.float64),
        lumo=np.array([m.lumo_energy for m in mols], dtype=np.float64),
        s1=np.array([m.s1_energy for m in mols], dtype=np.float64),
        t1=np.array([m.t1_energy for m in mols], dtype=np.float64),
    )


def _events(events, dims):
    return (np.array([site_index(e.initial, dims) for e in events], dtype=np.int32),
            np.array([site_index(e.final, dims) for e in events], dtype=np.int32),
            np.array([e.tau for e in events], dtype=np.float64))


def state_arrays(lat, prefix=""):
    dims = (lat._dimension.x, lat._dimension.y, lat._dimension.z)
    out = {}
    out[prefix + "e_pos"] = np.array([site_index(p, dims) for p in lat._electrons_locations], dtype=np.int32)
    out[prefix + "h_pos"] = np.array([site_index(p, dims) for p in lat._holes_locations], dtype=np.int32)
    out[prefix + "x_pos"] = np.array([site_index(p, dims) for p in lat._excitons_locations], dtype=np.int32)
    for name, lst in (("me", lat._move_electron_events), ("mh", lat._move_hole_events)):
        i, f, t = _events(lst, dims)
        out[prefix + name + "_init"], out[prefix + name + "_final"], out[prefix + name + "_tau"] = i, f, t
    special = []
    for lst in (lat._decay_events, lat._binding_events, lat._capture_events):
        for e in lst:
            special.append((e.kind, e.particule, site_index(e.initial, dims), e.tau))
    out[prefix + "special"] = np.array(special, dtype=np.float64).reshape(-1, 4)
    out[prefix + "tallies"] = np.array(
        [lat._emission, lat._recombination, lat._injection, lat._step], dtype=np.int64)
    return out


def run_reference(dims, props, charges, seed, n_steps, trajectory=0, realisation=0,
                  electric_field=0.1, distance=1, replay_traj=None, replay_spin=None):
    """Build one reference ``Lattice`` under injection, step it ``n_steps`` times and return a
    dict of numpy arrays: lattice, initial state, per-step trace, final state."""
    reseau, _, _, _ = import_reference()
    out = {}
    with injected(seed, trajectory, realisation, replay_traj, replay_spin) as ctx:
        lat = reseau.Lattice(tuple(dims), tuple(props), electric_field=electric_field,
                             charges=charges, charge_tranfer_distance=distance)
        d = (lat._dimension.x, lat._dimension.y, lat._dimension.z)
        out.update(lattice_arrays(lat))
        out.update(state_arrays(lat, "init_"))
        out["init_draws"] = np.array([ctx.traj.count, ctx.spin.count, ctx.species.count], dtype=np.int64)
        kind, part, ini, fin, tau, draws, spins = [], [], [], [], [], [], []
        error = ""
        for _ in range(n_steps):
            lat._step += 1                                   # reseau.py:582
            try:
                running = lat._first_reaction_method()       # reseau.py:586
            except (ZeroDivisionError, ValueError) as exc:   # reseau.py:587 / min() of empty (Q9)
                error = type(exc).__name__
                break
            if not running:                                  # reseau.py:590
                break
            ev = lat._cache[-1]
            kind.append(ev.kind)
            part.append(ev.particule)
            ini.append(site_index(ev.initial, d))
            fin.append(site_index(ev.final, d))
            tau.append(ev.tau)
            draws.append(ctx.traj.count)
            spins.append(ctx.spin.count)
        out["trace_kind"] = np.array(kind, dtype=np.uint8)
        out["trace_particule"] = np.array(part, dtype=np.uint8)
        out["trace_initial"] = np.array(ini, dtype=np.int32)
        out["trace_final"] = np.array(fin, dtype=np.int32)
        out["trace_tau"] = np.array(tau, dtype=np.float64)
        out["trace_draws"] = np.array(draws, dtype=np.int64)
        out["trace_spins"] = np.array(spins, dtype=np.int64)
        out.update(state_arrays(lat, "final_"))
        out["final_draws"] = np.array([ctx.traj.count, ctx.spin.count, ctx.species.count], dtype=np.int64)
        out["error"] = np.array(error)
        out["config"] = np.array([d[0], d[1], d[2], charges, distance, n_steps, trajectory, realisation],
                                 dtype=np.int64)
        out["props"] = np.array([lat._proportions.host, lat._proportions.tadf, lat._proportions.fluo],
                                dtype=np.float64)
        out["field"] = np.array(electric_field, dtype=np.float64)
        out["seed"] = np.array(seed, dtype=np.uint64)
    return out
