#!/usr/bin/env python
"""Freeze golden vectors by running the UNMODIFIED reference under injected Philox uniforms.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Writes ``tests/golden/lat_*.npz`` (one per lattice realisation: species + the four site
energies, exactly as the reference's ``Host``/``TADF``/``Fluorescent`` constructors produced
them) and ``tests/golden/traj_*.npz`` (initial charge/event state, the per-step trace
``(kind, particule, initial, final, tau, draws)`` read from ``Lattice._cache[-1]``, and the final
state and tallies).  The reference has no tests or fixtures of its own, so these
files are the parity pin for ``oracle/frm_oracle.c`` and for the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_harness as rh  # noqa: E402

DEFAULT_PROPS = (0.84, 0.15, 0.01)          # OLED.py:7
LATTICE_KEYS = ("species", "homo", "lumo", "s1", "t1")

# name, dims, props, charges, seed, steps, extra kwargs
CASES = [
    # the reference driver's own configuration (OLED.py:6-10), golden seeds 1..4
    ("d0_s1", (10, 10, 5), DEFAULT_PROPS, 4, 1, 3000, {}),
    ("d0_s2", (10, 10, 5), DEFAULT_PROPS, 4, 2, 3000, {}),
    ("d0_s3", (10, 10, 5), DEFAULT_PROPS, 4, 3, 3000, {}),
    ("d0_s4", (10, 10, 5), DEFAULT_PROPS, 4, 4, 3000, {}),
    # BASELINE config 1: 20^3, one electron + one hole in flight; same lattice reused at C=10
    ("c1_s1", (20, 20, 20), DEFAULT_PROPS, 1, 1, 3000, {}),
    ("c1_s2", (20, 20, 20), DEFAULT_PROPS, 1, 2, 3000, {}),
    ("c10_s1", (20, 20, 20), DEFAULT_PROPS, 10, 1, 2000, {}),
    ("c10_s2", (20, 20, 20), DEFAULT_PROPS, 10, 2, 2000, {}),
    # second trajectory on the same realisation (trajectory id != realisation id)
    ("c1_s1_t7", (20, 20, 20), DEFAULT_PROPS, 1, 1, 1500, {"trajectory": 7}),
    # ragged / minimal shapes: every site touches a wrap seam or a z face
    ("tiny333_s5", (3, 3, 3), (0.5, 0.3, 0.2), 2, 5, 1500, {}),
    ("rag456_s6", (4, 5, 6), (0.5, 0.3, 0.2), 3, 6, 2000, {}),
    ("rag734_s7", (7, 3, 4), (0.6, 0.3, 0.1), 3, 7, 2000, {}),
    # thin slab, many charges: blocked candidates, frozen charges (Q4), dense Coulomb sums
    ("dense_s8", (6, 6, 3), (0.6, 0.3, 0.1), 12, 8, 2000, {}),
    # Q1: proportions that do not sum to 1.0 are replaced by dimension/sum(dimension)
    ("q1_s9", (6, 6, 6), (0.5, 0.3, 0.3), 2, 9, 1000, {}),
    # field variants (sign convention Q8; zero field)
    ("field0_s10", (8, 8, 6), DEFAULT_PROPS, 3, 10, 1500, {"electric_field": 0.0}),
    ("field3_s11", (8, 8, 6), DEFAULT_PROPS, 3, 11, 1500, {"electric_field": 0.3}),
    # non-zero realisation id
    ("real3_s12", (8, 8, 8), DEFAULT_PROPS, 2, 12, 1500, {"realisation": 3, "trajectory": 5}),
    # 6 and 8 charges of each type: the upper half of the thread-per-trajectory kernel's range (hf_frm_small.cuh)
    ("c6_s13", (9, 8, 5), DEFAULT_PROPS, 6, 13, 1500, {}),
    ("c8_s14", (12, 10, 4), (0.7, 0.2, 0.1), 8, 14, 1500, {}),
    # 20 charges of each type: the warp kernel's screened candidate loop with the lanes sharing the Coulomb sums
    ("c20_s15", (12, 12, 5), DEFAULT_PROPS, 20, 15, 800, {}),
    # charges = 1 on lattices so small that the two carriers meet all the time: the thread-per-trajectory kernel's
    # bound / decay / capture / re-injection branches (hf_frm_c1.cuh) against the reference itself
    # (reseau.py:501-508, 529-561) — hundreds of recombinations and captures per case
    ("c1_tiny333_s21", (3, 3, 3), (0.5, 0.3, 0.2), 1, 21, 3000, {}),
    ("c1_rag443_s22", (4, 4, 3), (0.5, 0.3, 0.2), 1, 22, 3000, {}),
    ("c1_rag554_s23", (5, 5, 4), (0.6, 0.3, 0.1), 1, 23, 3000, {"trajectory": 3}),
    # charge_tranfer_distance > 1: _born_von_karman's table as the reference builds it (reseau.py:190-205, quirk Q3):
    # 124 neighbours in the bulk, asymmetric / duplicated entries near the seams, lattices thinner than 2d+1
    ("dist2_s31", (10, 10, 6), DEFAULT_PROPS, 4, 31, 1500, {"distance": 2}),
    ("dist2_thin_s32", (5, 4, 4), (0.5, 0.3, 0.2), 2, 32, 1200, {"distance": 2}),
    ("dist3_s33", (9, 8, 7), DEFAULT_PROPS, 3, 33, 800, {"distance": 3}),
    ("dist2_c1_s34", (8, 8, 5), DEFAULT_PROPS, 1, 34, 1500, {"distance": 2}),
]


def replay_case():
    """Crafted uniforms on the TRAJ stream: exercises ``_randbelow`` rejection
    (``random.py:264-269``: r >= limit is redrawn) and u = 0.0 (=> rng = 1.0, tau = -0.0)."""
    from oracle import philox as px
    n = 60000
    u = np.array([px.uniform(99, px.STREAM_TRAJ, 0, d) for d in range(n)], dtype=np.float64)
    top = 1.0 - 2.0 ** -53                     # largest double below 1: always >= limit unless n | 2^53
    u[0] = top                                  # first draw of the electron injection sample -> rejected
    u[5] = top
    u[40] = 0.0                                 # a candidate with rng = 1.0 -> tau = -0.0 wins its min()
    u[200:4000:97] = top                        # sprinkle rejections over re-injection choices
    u[300:4000:131] = 0.0
    return u


def reference_statistics(n_runs=400, dims=(3, 3, 3), props=(0.5, 0.3, 0.2), charges=2, steps=1500, name="stats_tiny"):
    """Tallies of n_runs UNMODIFIED reference runs on their own Mersenne-Twister streams (the only
    change is that each `Random()` gets a deterministic seed so the file is reproducible): the
    statistical half of the parity contract (3 sigma on IQE and channel fractions)."""
    import itertools
    import random
    reseau, molecule, _, _ = rh.import_reference()
    counter = itertools.count(1)
    saved = (reseau.Random, molecule.Random)
    reseau.Random = molecule.Random = lambda: random.Random(0x5EED0000 + next(counter))
    em, rec, inj, stp = [], [], [], []
    try:
        for _ in range(n_runs):
            lat = reseau.Lattice(dims, props, charges=charges)
            lat.operations(steps)
            em.append(lat._emission); rec.append(lat._recombination); inj.append(lat._injection); stp.append(lat._step)
    finally:
        reseau.Random, molecule.Random = saved
    np.savez_compressed(os.path.join(HERE, name + ".npz"), dims=np.array(dims), props=np.array(props),
                        charges=np.array(charges), steps=np.array(steps), emission=np.array(em),
                        recombination=np.array(rec), injection=np.array(inj), step=np.array(stp))
    print(f"{name}: runs={n_runs} emission/run={np.mean(em):.2f} recombination/run={np.mean(rec):.2f} "
          f"photons/recombination={np.sum(em) / np.sum(rec):.4f}")


def _channel_run(args):
    """One UNMODIFIED reference run of the OLED.py device with its own Mersenne-Twister streams (seeded per run so
    that the file is reproducible).  The channel tallies the reference does not keep — which species the exciton
    decayed on and whether it emitted — are observed by wrapping the *instance's* ``_decay`` (reseau.py:472-476);
    ``operations`` itself runs untouched, swallowed ZeroDivisionError and stale IQE included."""
    run, dims, props, charges, steps = args
    import itertools
    import random
    reseau, molecule, _, _ = rh.import_reference()
    counter = itertools.count(1)
    reseau.Random = molecule.Random = lambda: random.Random(0x5EED0000 + run * 100003 + next(counter))
    import io
    import contextlib
    code = {molecule.Host: 0, molecule.TADF: 1, molecule.Fluorescent: 2}
    lat = reseau.Lattice(dims, props, charges=charges)
    ch = [0, 0, 0, 0, 0, 0]                       # recombinations on host / TADF / fluorescent, emissions on the same
    inner = lat._decay
    def observed(position):
        sp = code[lat._get_molecule_type(position)]
        before = lat._emission
        inner(position)
        ch[sp] += 1
        ch[3 + sp] += lat._emission - before
    lat._decay = observed
    with contextlib.redirect_stdout(io.StringIO()):      # the _cache dump of a swallowed ZeroDivisionError
        lat.operations(steps)
    return (lat._emission, lat._recombination, lat._injection, lat._step, lat._IQE, *ch)


def reference_channel_statistics(n_runs=100_000, dims=(10, 10, 5), props=DEFAULT_PROPS, charges=4, steps=100,
                                 name="stats_d0_1e5", procs=None):
    """The north_star's statistical gate: tallies AND channel counts of 10^5 unmodified runs of the reference
    driver's own device (OLED.py:6-11: 10x10x5, charges = 4, operations(100))."""
    import multiprocessing as mp
    procs = procs or os.cpu_count()
    with mp.Pool(procs) as pool:
        rows = pool.map(_channel_run, [(r, dims, props, charges, steps) for r in range(n_runs)], chunksize=250)
    a = np.array([r[:4] + r[5:] for r in rows], dtype=np.int32)
    iqe = np.array([r[4] for r in rows], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), dims=np.array(dims), props=np.array(props),
                        charges=np.array(charges), steps=np.array(steps),
                        emission=a[:, 0].astype(np.int16), recombination=a[:, 1].astype(np.int16),
                        injection=a[:, 2].astype(np.int16), step=a[:, 3].astype(np.int16), iqe=iqe.astype(np.float32),
                        recomb_host=a[:, 4].astype(np.int8), recomb_tadf=a[:, 5].astype(np.int8), recomb_fluo=a[:, 6].astype(np.int8),
                        emit_host=a[:, 7].astype(np.int8), emit_tadf=a[:, 8].astype(np.int8), emit_fluo=a[:, 9].astype(np.int8))
    print(f"{name}: runs={n_runs} recombination/run={a[:, 1].mean():.4f} emission/run={a[:, 0].mean():.4f} "
          f"recomb host/tadf/fluo={a[:, 4].sum()}/{a[:, 5].sum()}/{a[:, 6].sum()} "
          f"emit host/tadf/fluo={a[:, 7].sum()}/{a[:, 8].sum()}/{a[:, 9].sum()} early stops={(a[:, 3] < steps).sum()}")


def main():
    os.makedirs(HERE, exist_ok=True)
    if "--stats-only" in sys.argv:
        reference_statistics()
        return
    if "--stats-d0-1e5" in sys.argv:             # 10^5 runs of OLED.py's device as OLED.py runs it (100 operations), with channels
        reference_channel_statistics()
        return
    if "--stats-d0" in sys.argv:                 # the reference driver's own device (OLED.py:6-10), 2000 operations
        reference_statistics(n_runs=150, dims=(10, 10, 5), props=DEFAULT_PROPS, charges=4, steps=2000, name="stats_d0")
        return
    written = {}
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    if only:                                     # add cases without regenerating the committed ones
        for name, dims, props, charges, seed, steps, kw in CASES:
            if name in only:
                out = rh.run_reference(dims, props, charges, seed, steps, **kw)
                _write(name, out, dims, seed, kw.get("realisation", 0), written)
        return
    for name, dims, props, charges, seed, steps, kw in CASES:
        out = rh.run_reference(dims, props, charges, seed, steps, **kw)
        _write(name, out, dims, seed, kw.get("realisation", 0), written)
    u = replay_case()
    out = rh.run_reference((6, 5, 4), (0.6, 0.3, 0.1), 3, 99, 1500, replay_traj=u)
    used = int(out["final_draws"][0])
    out["replay_traj"] = u[:used + 8]
    _write("replay_s99", out, (6, 5, 4), 99, 0, written)
    reference_statistics()


def _write(name, out, dims, seed, real, written):
    lat_name = f"lat_{dims[0]}x{dims[1]}x{dims[2]}_s{seed}_r{real}"
    lat = {k: out.pop(k) for k in LATTICE_KEYS}
    if lat_name in written:
        for k in LATTICE_KEYS:                   # same (seed, realisation) => identical lattice
            assert np.array_equal(lat[k], written[lat_name][k]), (name, k)
    else:
        np.savez_compressed(os.path.join(HERE, lat_name + ".npz"), **lat)
        written[lat_name] = lat
    out["lattice"] = np.array(lat_name)
    np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), **out)
    t = out["final_tallies"]
    print(f"{name:12s} steps={len(out['trace_kind']):5d} emission={t[0]} recomb={t[1]} inj={t[2]} "
          f"draws={out['final_draws'].tolist()} error={str(out['error'])!r}")


if __name__ == "__main__":
    main()
