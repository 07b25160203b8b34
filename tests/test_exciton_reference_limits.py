"""The exciton extensions against the only exciton facts the reference defines.

The reference has no exciton transport, so the exciton kernels are specified by this
repository's own oracles (parity unpinned).  It does define three facts, and in the limits where nothing else
acts they pin both the specifications (CPU, here) and the CUDA kernels (GPU, below) to the reference:

  F1  an exciton is born a singlet with probability 0.25, else a triplet       (molecule.py:88-96)
  F2  a decay emits a photon iff the exciton is a singlet on a TADF or a fluorescent molecule; never on a host
                                                                                (molecule.py:185-199, 283-297, 387-400)
  F3  intersystem crossing flips singlet <-> triplet                            (molecule.py:299-307)

Degenerate limit (transfer, ISC and RISC rates 0, every decay channel certain): photons per exciton decayed on a
species must equal what the UNMODIFIED reference produced on that species — tests/golden/stats_d0_1e5.npz, 10^5
runs of OLED.py's device with the species of every decay observed (make_golden.py --stats-d0-1e5) — within 3
sigma, and photons per exciton overall must be 0.25 * P(site is TADF or fluorescent).
ISC-only limit (S1 = T1 everywhere, so the flip carries no Boltzmann factor): the flip alternates the spin and the
photon yield follows the two-state chain's closed form.
"""
import os

import numpy as np
import pytest

from oracle import exciton_oracle as xo
from oracle import frm_oracle as fo
from tests.conftest import GOLDEN_DIR

K = 1e8
PROPS = (0.84, 0.15, 0.01)


def degenerate_params(cutoff=2.0, k_isc=0.0, k_risc=0.0, k_t=K):
    """Nothing but decay: singlets radiate on TADF / fluorescent molecules and are quenched on hosts, triplets are
    quenched everywhere (F2 turned into rates)."""
    z = np.zeros((3, 3))
    return dict(xo.default_params(), cutoff=cutoff, forster_radius=z, dexter_prefactor=z,
                k_isc=np.array([0.0, k_isc, 0.0]), k_risc=np.array([0.0, k_risc, 0.0]),
                k_rad=np.array([[0.0, 0.0], [K, 0.0], [K, 0.0]]),
                k_nonrad=np.array([[K, k_t], [0.0, k_t], [0.0, k_t]]))


def reference_yields():
    ref = np.load(os.path.join(GOLDEN_DIR, "stats_d0_1e5.npz"))
    out = {}
    for sp in ("host", "tadf", "fluo"):
        out[sp] = (float(ref["emit_" + sp].sum()), float(ref["recomb_" + sp].sum()))
    return out


def two_sample(k1, n1, k2, n2):
    """|p1 - p2| in units of the pooled binomial sigma."""
    pool = (k1 + k2) / (n1 + n2)
    sigma = np.sqrt(max(pool * (1 - pool), 1e-300) * (1 / n1 + 1 / n2))
    return abs(k1 / n1 - k2 / n2) / sigma


def check_degenerate(t, nonhost_fraction, decayed):
    """t: exciton tallies (HF_XT_* names) of a degenerate-limit run."""
    assert t["forster"] == t["dexter"] == t["isc"] == t["risc"] == 0
    assert t["rad_host"] == 0                                                           # F2: hosts never emit
    ref = reference_yields()
    assert ref["host"][0] == 0 and ref["host"][1] > 1000
    for sp in ("tadf", "fluo"):
        n = t["rad_" + sp] + t["nonrad_s_" + sp] + t["nonrad_t_" + sp]
        assert t["nonrad_s_" + sp] == 0                                                 # F2: a singlet there always emits
        assert n > 1000
        z = two_sample(t["rad_" + sp], n, *ref[sp])                                     # vs the unmodified reference
        assert z < 3, (sp, t["rad_" + sp] / n, ref[sp][0] / ref[sp][1], z)
        assert abs(t["rad_" + sp] / n - 0.25) < 3 * np.sqrt(0.25 * 0.75 / n)            # F1
    n_host = t["nonrad_s_host"] + t["nonrad_t_host"]
    assert abs(t["nonrad_s_host"] / n_host - 0.25) < 3 * np.sqrt(0.25 * 0.75 / n_host)  # F1 on hosts
    photons = t["rad_tadf"] + t["rad_fluo"]
    p = 0.25 * nonhost_fraction                                                         # 0.25 * P(TADF or fluorescent)
    assert sum(t[k] for k in t if k.startswith(("rad_", "nonrad_"))) == decayed
    assert abs(photons / decayed - p) < 3 * np.sqrt(p * (1 - p) / decayed), (photons / decayed, p)


def isc_chain_yield(kappa, k_s, k_t):
    """Photon probability of a TADF exciton that only flips (rate kappa both ways, F3), radiates as a singlet
    (k_s) and is quenched as a triplet (k_t), born a singlet with probability 0.25 (F1)."""
    p_s = (k_s / (k_s + kappa)) / (1.0 - kappa * kappa / ((k_s + kappa) * (k_t + kappa)))
    p_t = kappa / (kappa + k_t) * p_s
    return 0.25 * p_s + 0.75 * p_t, p_s, p_t


def _lattice(dims, seed, flat_st=False):
    a = fo.OracleLattice.generate(dims, PROPS, seed).arrays()
    if flat_st:
        a["t1"] = a["s1"].copy()                                                        # S1 = T1: RISC carries no Boltzmann factor
    return a


def _nonhost_fraction(species, dims):
    interior = species.reshape(dims[2], dims[1], dims[0])[1:-1]
    return float((interior != 0).mean())


def test_specification_degenerate_limit_equals_the_reference_facts():
    dims, seed = (16, 16, 8), 31
    a = _lattice(dims, seed)
    ws1, wt1 = xo.weights(a["s1"], a["t1"], a["species"])
    o = xo.ExcitonOracle(dims, a["species"], ws1, wt1, degenerate_params())
    n_traj, n_events = 250_000, 4
    t = o.run_batch(seed, 0, n_traj, n_events, n_threads=os.cpu_count() or 1)
    assert t["events"] == n_traj * n_events                                             # every event is a decay
    check_degenerate(t, _nonhost_fraction(a["species"], dims), n_traj * n_events)


def test_specification_isc_only_limit_flips_like_the_reference():
    dims, seed, kappa, k_t = (12, 12, 6), 32, 3e8, 5e7
    a = _lattice(dims, seed, flat_st=True)
    ws1, wt1 = xo.weights(a["s1"], a["t1"], a["species"])
    o = xo.ExcitonOracle(dims, a["species"], ws1, wt1, degenerate_params(k_isc=kappa, k_risc=kappa, k_t=k_t))
    # F3 event by event: on a TADF site every ISC / RISC event flips the spin, nothing else changes it
    flips = 0
    for traj in range(200):
        tr = o.run(seed, traj, 60)["trace"]
        for kind, spin_after, prev in zip(tr["kind"][1:], tr["spin"][1:], tr["spin"][:-1]):
            if kind == 3:                                                               # EVENTS["ISC"] (event.py:46)
                assert {int(prev), int(spin_after)} == {1, 3}
                flips += 1
    assert flips > 200
    # long trajectories: an exciton still alive when its trajectory ends is not counted, which biases the yield of a
    # short run towards short-lived excitons; 2000 events per trajectory push that bias far below the 3 sigma band
    t = o.run_batch(seed, 0, 4000, 2000, n_threads=os.cpu_count() or 1)
    n = t["rad_tadf"] + t["nonrad_s_tadf"] + t["nonrad_t_tadf"]
    y, _, _ = isc_chain_yield(kappa, K, k_t)
    assert n > 50_000 and t["isc"] > 0 and t["risc"] > 0
    assert abs(t["rad_tadf"] / n - y) < 3 * np.sqrt(y * (1 - y) / n), (t["rad_tadf"] / n, y)
    # fluorescent molecules and hosts do not cross (k_isc = 0 there): still F1 x F2
    nf = t["rad_fluo"] + t["nonrad_t_fluo"]
    assert abs(t["rad_fluo"] / nf - 0.25) < 3 * np.sqrt(0.25 * 0.75 / nf)


def test_dense_specification_degenerate_limit_equals_the_reference_facts():
    dims, seed = (16, 16, 8), 33
    a = _lattice(dims, seed)
    ws1, wt1 = xo.weights(a["s1"], a["t1"], a["species"])
    o = xo.DenseExcitonOracle(dims, a["species"], ws1, wt1, degenerate_params(), annihilation=((0, 0), (0, 0)), p_tta=0.0)
    n_dev, n_exc, n_events = 2048, 32, 300
    t = o.run_devices(seed, 0, n_dev, n_exc, n_events, n_threads=os.cpu_count() or 1)
    assert t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"] + t["tta_up"] == 0
    decayed = n_dev * n_events
    t15 = {k: v for k, v in t.items() if not k.startswith(("ann_", "tta_"))}
    check_degenerate(t15, _nonhost_fraction(a["species"], dims), decayed)


# ---- the CUDA kernels in the same limits ---------------------------------------------------------------------

def _gpu_params(nat, **kw):
    return nat.x_params_from_dict(degenerate_params(**kw))


@pytest.mark.gpu
def test_gpu_exciton_kernel_degenerate_limit_equals_the_reference_facts():
    """hfx_run_kernel, 10^6 trajectories: > 2 x 10^6 excitons decayed."""
    from hyperfluorescence_b200 import _native as nat
    dims, B, n = (32, 32, 16), 1_000_000, 2
    sim = nat.Sim(dims, PROPS, 0.1, 1, 1, n_trajectories=B, seed=41)
    try:
        sim.generate_lattice()
        sim.x_configure(_gpu_params(nat))
        sim.x_spawn()
        sim.x_run(n)
        sim.sync()
        t = {k: v for k, v in sim.x_tallies().items() if not k.startswith("_")}
        species = sim.download_lattice(0)["species"]
    finally:
        sim.close()
    assert t["events"] == B * n
    check_degenerate(t, _nonhost_fraction(species, dims), B * n)


@pytest.mark.gpu
def test_gpu_exciton_kernel_isc_only_limit():
    from hyperfluorescence_b200 import _native as nat
    dims, B, n, kappa, k_t = (16, 16, 8), 8192, 2000, 3e8, 5e7     # long trajectories: see the CPU twin of this test
    a = _lattice(dims, 42, flat_st=True)
    sim = nat.Sim(dims, PROPS, 0.1, 1, 1, n_trajectories=B, seed=42)
    try:
        sim.upload_lattice(0, a["species"], a["homo"], a["lumo"], a["s1"], a["t1"])
        sim.x_configure(_gpu_params(nat, k_isc=kappa, k_risc=kappa, k_t=k_t))
        sim.set_trace(0, 64, 100)
        sim.x_spawn()
        sim.x_run(n)
        sim.sync()
        t = sim.x_tallies()
        traces = [sim.x_trace(b, 100) for b in range(64)]
    finally:
        sim.close()
    n_tadf = t["rad_tadf"] + t["nonrad_s_tadf"] + t["nonrad_t_tadf"]
    y, _, _ = isc_chain_yield(kappa, K, k_t)
    assert n_tadf > 100_000 and t["isc"] > 0 and t["risc"] > 0
    assert abs(t["rad_tadf"] / n_tadf - y) < 3 * np.sqrt(y * (1 - y) / n_tadf), (t["rad_tadf"] / n_tadf, y)
    assert t["forster"] == t["dexter"] == 0 and t["rad_host"] == 0
    assert sum(int((tr["kind"] == 3).sum()) for tr in traces) > 20                      # flips were traced too


@pytest.mark.gpu
def test_gpu_dense_kernel_degenerate_limit_equals_the_reference_facts():
    """hfxd_run_kernel: 4096 devices x 64 excitons, 256 events per device: > 10^6 excitons decayed."""
    from hyperfluorescence_b200 import _native as nat
    dims, B, n_exc, n = (32, 32, 16), 4096, 64, 256
    sim = nat.Sim(dims, PROPS, 0.1, 1, 1, n_trajectories=B, seed=43)
    try:
        sim.generate_lattice()
        sim.x_configure(_gpu_params(nat))
        sim.xd_configure(n_exc, annihilation=((0.0, 0.0), (0.0, 0.0)), p_tta=0.0)
        sim.xd_spawn()
        sim.xd_run(n)
        sim.sync()
        t = sim.xd_tallies()
        species = sim.download_lattice(0)["species"]
    finally:
        sim.close()
    assert t["ann_ss"] + t["ann_st"] + t["ann_ts"] + t["ann_tt"] + t["tta_up"] == 0
    t15 = {k: v for k, v in t.items() if not k.startswith(("ann_", "tta_"))}
    check_degenerate(t15, _nonhost_fraction(species, dims), B * n)
