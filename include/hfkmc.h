/*
 * hfkmc.h — C-ABI of libhfkmc.so, the B200 (sm_100a) implementation of hyperfluorescence's
 * kinetic-Monte-Carlo hot path.
 *
 * The reference (Pehnny/hyperfluorescence) is pure Python and has no FFI of its own; the
 * boundary it offers is the object surface of reseau.Lattice.  Every entry point below cites
 * the reference interface it replaces (file:line into the reference tree).  The Python façade
 * hyperfluorescence_b200/reseau.py binds these with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; all array arguments are HOST buffers owned by the caller unless the
 *     name says `_device`; outputs are copied into caller memory, no pointer into device state
 *     ever escapes;
 *   - every function returns an hf_status (0 = ok); hf_last_error() gives the message of the
 *     last failure on the calling thread; no C++ exception crosses the boundary;
 *   - a handle owns its device buffers and one CUDA stream; a handle is not thread-safe,
 *     different handles are independent;
 *   - hf_run / hf_inject / hf_generate_lattice are asynchronous with respect to the host;
 *     every hf_read_* synchronises the handle's stream first;
 *   - there is NO CPU fallback: hf_create fails with HF_ERR_NO_DEVICE when no CUDA device of
 *     compute capability 10.x is visible.
 *
 * Site indices are flat, site = (z*Y + y)*X + x (the order in which reseau.py:173 creates the
 * molecules).  Event / particle / exciton / species codes are the reference's
 * (event.py:43-57, molecule.py:16-21).
 */
#ifndef HFKMC_H
#define HFKMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HF_ABI_VERSION 2

typedef struct hf_sim hf_sim;

typedef enum hf_status {
    HF_OK = 0,
    HF_ERR_BAD_ARG = 1,       /* TypeError / IndexError / ValueError of reseau.py:126-139, 212-213 */
    HF_ERR_CUDA = 2,          /* a CUDA runtime call failed; see hf_last_error()                   */
    HF_ERR_NO_DEVICE = 3,     /* no sm_100 device: the product has no CPU path                     */
    HF_ERR_UNSUPPORTED = 4,   /* capacity limits (dimensions > 1023, charge_tranfer_distance > 8)  */
    HF_ERR_STATE = 5          /* call order (e.g. hf_run before a lattice exists)                  */
} hf_status;

/* Why a trajectory stopped inside hf_run (hf_traj_info.status). */
typedef enum hf_traj_status {
    HF_TRAJ_OK = 0,
    HF_TRAJ_NO_EVENTS = 1,        /* _first_reaction_method returned False   reseau.py:488, 590  */
    HF_TRAJ_ZERO_DIVISION = 2,    /* ZeroDivisionError swallowed by operations reseau.py:587-589 */
    HF_TRAJ_VALUE_ERROR = 3,      /* min() of no candidates reseau.py:387,396 / list.remove miss */
    HF_TRAJ_REPLAY_EXHAUSTED = 4, /* hf_set_replay buffer too short                              */
    HF_TRAJ_CAPACITY = 5          /* internal flag-set capacity exceeded (fails loudly)          */
} hf_traj_status;

/* event.py:43-51 */
enum { HF_EV_MOVE = 1, HF_EV_BOUND = 2, HF_EV_ISC = 3, HF_EV_FORSTER = 4, HF_EV_DECAY = 5, HF_EV_UNBOUND = 6, HF_EV_CAPTURE = 7 };
/* event.py:53-57 */
enum { HF_P_ELECTRON = 1, HF_P_HOLE = 2, HF_P_EXCITON = 3 };
/* molecule.py:16-21 */
enum { HF_X_NONE = 0, HF_X_SINGLET = 1, HF_X_DOUBLET = 2, HF_X_TRIPLET = 3 };
/* reseau.py:175-182 (_molecule_type codes) */
enum { HF_SP_HOST = 0, HF_SP_TADF = 1, HF_SP_FLUO = 2 };

/* hf_config.flags */
#define HF_FLAG_NONE 0u
/* Step trajectories with charges <= 12 with the generic warp-per-trajectory kernel instead of the
 * thread-per-trajectory ones (cross-checks; all give identical results). */
#define HF_FLAG_WARP_KERNEL 1u

/* Replaces the arguments of Lattice.__init__ (reseau.py:109-111) plus the batch shape. */
typedef struct hf_config {
    int32_t dims[3];             /* dimension (x, y, z); z >= 3 (reseau.py:131)                      */
    double props[3];             /* (host, tadf, fluo) AFTER quirk Q1 (reseau.py:143-145)            */
    double electric_field;       /* eV/nm, default 0.1 (reseau.py:110)                               */
    int32_t charges;             /* electrons = holes = charges (reseau.py:152)                      */
    int32_t distance;            /* charge_tranfer_distance 1..8: reseau.py:190-205 verbatim (Q3)    */
    int64_t n_trajectories;      /* independent Lattice instances ("devices") held by this handle    */
    int32_t n_realisations;      /* lattice realisations; must divide n_trajectories; trajectory i   */
                                 /* runs on realisation i / (n_trajectories / n_realisations)        */
    uint32_t first_trajectory;   /* global ids (Philox stream idents) of local trajectory 0 ...      */
    uint32_t first_realisation;  /* ... and local realisation 0: results do not depend on sharding   */
    uint64_t seed;               /* Philox key                                                       */
    uint32_t flags;
    int32_t device;              /* CUDA device ordinal                                              */
} hf_config;

/* One fired event: the tuple Lattice._cache holds (reseau.py:124, 489-490; event.py:114-139). */
typedef struct hf_event {
    int32_t initial;             /* site index of Event.initial                                      */
    int32_t final_;              /* site index of Event.final                                        */
    double tau;                  /* Event.tau                                                        */
    uint8_t kind;                /* HF_EV_*                                                          */
    uint8_t particule;           /* HF_P_*                                                           */
    uint8_t pad[2];
    uint32_t spins;              /* SPIN-stream draws consumed after the step (trace only)           */
    uint64_t draws;              /* TRAJ-stream draws consumed after the step (trace only)           */
} hf_event;

/* Scalar state of one trajectory: the counters of reseau.py:118-123 and list lengths. */
typedef struct hf_traj_info {
    int32_t n_electrons, n_holes;      /* len(_electrons_locations), len(_holes_locations)          */
    int32_t n_eflags, n_hflags;        /* sites whose Molecule.electron / .hole flag is True        */
    int32_t n_move_e, n_move_h;        /* len(_move_electron_events), len(_move_hole_events)        */
    int32_t special_kind;              /* pending bound / decay / capture event (0 = none)          */
    int32_t special_particule;
    int32_t special_site;
    int32_t exciton_state;             /* HF_X_* of the exciton awaiting decay                      */
    int32_t status;                    /* hf_traj_status                                            */
    int32_t cache_count;               /* events pushed into _cache so far, saturating at 10        */
    uint64_t draws, spins;             /* uniforms consumed on the TRAJ / SPIN streams              */
    uint64_t emission, recombination, injection, steps;   /* _emission, _recombination, _injection, _step */
    uint64_t fired;                    /* events actually fired (= trace length)                    */
} hf_traj_info;

/* Box totals, hf_read_tallies / hf_reduce_tallies_device (u64 each). */
enum {
    HF_T_EMISSION = 0, HF_T_RECOMBINATION = 1, HF_T_INJECTION = 2, HF_T_STEPS = 3, HF_T_FIRED = 4,
    HF_T_MOVE_E = 5, HF_T_MOVE_H = 6, HF_T_BOUND = 7, HF_T_DECAY = 8, HF_T_CAPTURE_E = 9, HF_T_CAPTURE_H = 10,
    HF_T_RECOMB_HOST = 11, HF_T_RECOMB_TADF = 12, HF_T_RECOMB_FLUO = 13,
    HF_T_EMIT_TADF = 14, HF_T_EMIT_FLUO = 15,
    HF_T_STOPPED = 16,           /* trajectories whose status != HF_TRAJ_OK                          */
    HF_T_X_FORSTER = 17, HF_T_X_DEXTER = 18, HF_T_X_ISC = 19, HF_T_X_RISC = 20,   /* integrated excitons (hf_p_excitons)  */
    HF_N_TALLIES = 24
};

const char *hf_last_error(void);
int hf_abi_version(void);
/* Number of visible sm_100 devices (0 => hf_create will fail; never falls back to the CPU). */
int hf_device_count(void);

/* Lattice.__init__ argument validation and parameter block (reseau.py:109-152) + allocation of
 * the device state for config.n_trajectories trajectories.  No lattice exists yet. */
int hf_create(const hf_config *config, hf_sim **out);
int hf_destroy(hf_sim *sim);

/* Launch all work of this handle on `cuda_stream` (a cudaStream_t) instead of its own stream. */
int hf_set_stream(hf_sim *sim, void *cuda_stream);
int hf_sync(hf_sim *sim);

/* Lattice._lattice_creation (reseau.py:154-173) + the species constructors
 * (molecule.py:156-183, 252-281, 358-385), on the device, for every realisation, from the
 * Philox SPECIES / ENERGY streams (draw contract in DESIGN.md). */
int hf_generate_lattice(hf_sim *sim);
/* Same, with one composition per realisation: props[n_realisations][3] = (host, tadf, fluo) after
 * Q1.  This is the fitness kernel of the composition sweep the reference names as its goal
 * ("optimize the device through the molecules proportions", README.md:2): candidate r is
 * realisation r, its trajectories are the n_trajectories / n_realisations that run on it. */
int hf_generate_lattice_mixed(hf_sim *sim, const double *props);
/* Import a lattice built elsewhere (e.g. by the reference): replaces Lattice._grid's
 * molecules for one realisation.  s1 / t1 may be NULL. */
int hf_upload_lattice(hf_sim *sim, int32_t realisation, const uint8_t *species, const double *homo,
                      const double *lumo, const double *s1, const double *t1);
/* Read back what Lattice._grid would hold: type codes and the four energies of every site. */
int hf_download_lattice(hf_sim *sim, int32_t realisation, uint8_t *species, double *homo,
                        double *lumo, double *s1, double *t1);

/* Lattice._charges_injection + _events_creation + the tally initialisation
 * (reseau.py:116-124, 207-276) for every trajectory. */
int hf_inject(hf_sim *sim);

/* Replay mode: the TRAJ stream of local trajectory i reads u[i*n_per_trajectory + d] instead of
 * Philox (the "injected uniforms" of the parity contract; replaces Lattice._seed.random(),
 * reseau.py:284, 316).  u == NULL switches back to Philox.  Call before hf_inject. */
int hf_set_replay(hf_sim *sim, const double *u, int64_t n_per_trajectory);

/* Lattice.operations(stop) (reseau.py:580-595) for every trajectory of the handle. */
int hf_run(hf_sim *sim, int64_t stop);

/* Sums over the handle's trajectories, out[HF_N_TALLIES]: _emission, _recombination, _injection,
 * _step (reseau.py:118-122) and the per-kind / per-species channel counts. */
int hf_read_tallies(hf_sim *sim, uint64_t *out);
/* Same totals written to a DEVICE buffer of HF_N_TALLIES u64 on the handle's stream (for an
 * NCCL all-reduce across ranks without a host round trip). */
int hf_reduce_tallies_device(hf_sim *sim, void *out_device);

/* Per-realisation sums, out[n_realisations][5] = _emission, _recombination, _injection, _step,
 * events fired (reseau.py:118-122) over the trajectories of each realisation. */
int hf_read_realisation_tallies(hf_sim *sim, uint64_t *out);

/* Per-trajectory scalar state (reseau.py:118-123). */
int hf_read_info(hf_sim *sim, int64_t trajectory, hf_traj_info *out);
/* The same for trajectories [first, first + n) in one copy. */
int hf_read_infos(hf_sim *sim, int64_t first, int64_t n, hf_traj_info *out);
/* Lattice._electrons_locations / _holes_locations / _excitons_locations in list order
 * (which = 0 / 1 / 2; reseau.py:208-210, 617-621).  *n in: capacity, out: length. */
int hf_read_positions(hf_sim *sim, int64_t trajectory, int32_t which, int32_t *sites, int32_t *n);
/* Integrated exciton dynamics — the slots the reference reserves and leaves empty: _move_exciton_events /
 * _isc_events / _decay_events (reseau.py:233-235), the `...` branches of the dispatcher (reseau.py:526-550, routing
 * comment 533), TADF.intersystem_crossing (molecule.py:299-307).  OWN MODEL beyond mode 1 (specification: the
 * "integrated excitons" block of oracle/frm_oracle.c).  mode 0: off, the reference as it is (the exciton decays at
 * tau = 0 on the site it formed on, reseau.py:529-533); 1: the same behaviour routed through the exciton slots and
 * lists (bit-identical to mode 0: the parity gate of the new code path); 2: the exciton formed by a `bound` event
 * hops (Forster / Dexter), crosses (ISC / RISC) and decays (radiatively or not) in the SAME First Reaction Method as
 * the carriers, with the parameters, tables and Boltzmann weights of hf_x_configure (call it first).  Changing the
 * mode needs a new hf_inject. */
int hf_p_excitons(hf_sim *sim, int32_t mode);
/* Exciton slots of one trajectory in creation order: site, spin (HF_X_*), and the pending event of each — kind
 * (HF_EV_FORSTER / HF_EV_MOVE transfer, HF_EV_ISC, HF_EV_DECAY), final site, residual tau, radiative flag: the
 * contents of _move_exciton_events / _isc_events / _decay_events.  *n in: capacity, out: count.  draws: uniforms
 * consumed on the trajectory's exciton stream; clock: simulated time (s).  Any output may be NULL. */
int hf_read_excitons(hf_sim *sim, int64_t trajectory, int32_t *site, int32_t *spin, int32_t *kind, int32_t *final_,
                     double *tau, int32_t *radiative, int32_t *n, uint64_t *draws, double *clock);
/* Sites whose Molecule.electron (which=0) / .hole (which=1) flag is True (molecule.py:69-70). */
int hf_read_flags(hf_sim *sim, int64_t trajectory, int32_t which, int32_t *sites, int32_t *n);
/* Lattice._move_electron_events (which=0) / _move_hole_events (which=1) in list order
 * (reseau.py:231-232). */
int hf_read_move_events(hf_sim *sim, int64_t trajectory, int32_t which, int32_t *initial,
                        int32_t *final_, double *tau, int32_t *n);
/* Lattice._cache (reseau.py:124): the last <= 10 fired events, oldest first. */
int hf_read_last_events(hf_sim *sim, int64_t trajectory, hf_event *out, int32_t *n);

/* Per-step trace for parity tests: records every fired event of local trajectories
 * [first, first + n) up to `capacity` events each, from the next hf_run on.  n = 0 disables. */
int hf_set_trace(hf_sim *sim, int64_t first, int64_t n, int64_t capacity);
int hf_read_trace(hf_sim *sim, int64_t trajectory, hf_event *out, int64_t *n);

/* Lattice._time_move_electron / _time_move_hole (reseau.py:283-292, 315-324) for hand-built
 * candidate hops of one trajectory in its CURRENT state, given the uniform each would draw:
 * tau[i] = -log(1 - u[i]) / k(initial[i] -> final[i]).  rate[i] receives k.  Known-answer tests. */
int hf_eval_hops(hf_sim *sim, int64_t trajectory, int32_t particule, int32_t n, const int32_t *initial,
                 const int32_t *final_, const double *u, double *tau, double *rate, int32_t *zero_division);

/* Timing helper for bench.py: TOTAL device time (ms) of the hf_run / hf_x_run / hf_xd_run launches issued
 * since the last call, measured with CUDA events on the handle's stream, and their count (the average is
 * total_ms / launches).  Resets both. */
int hf_run_timing(hf_sim *sim, double *total_ms, int64_t *launches);

/* =====================================================================================================
 * Exciton trajectories ("Path X").  The reference has NO implementation of these (parity unpinned):
 * it only provides the slots — event codes ISC / Forster / decay (event.py:43-51), the empty
 * _move_exciton_events / _isc_events lists (reseau.py:233-235), the `...` branches of the dispatcher
 * (reseau.py:526-550), TADF.intersystem_crossing (molecule.py:299-307), the singlet fraction
 * (molecule.py:93) and the per-site S1 / T1 energies (molecule.py:182-183).  The model is this
 * library's own; it is specified by oracle/exciton_oracle.c and DESIGN.md section 9.
 * One exciton per trajectory of the handle, on the handle's lattice realisations; rejection-free
 * (BKL) event selection; a decayed exciton is replaced by a new one at a random interior site.
 * ===================================================================================================== */

/* Rate-law parameters; species index 0 host, 1 TADF, 2 fluorescent; spin index 0 singlet, 1 triplet.
 * None of these numbers exists in the reference: hf_x_default_params documents the defaults. */
typedef struct hf_x_params {
    double cutoff;                    /* transfer cutoff radius in lattice constants (1 nm)              */
    double forster_radius[3][3];      /* R0[donor][acceptor] in nm: k = k_S(donor) (R0 / r)^6            */
    double dexter_prefactor[3][3];    /* k_D0[donor][acceptor] in 1/s: k = k_D0 exp(-2 r / L)           */
    double dexter_length;             /* L in nm                                                          */
    double k_isc[3];                  /* S -> T, 1/s                                                      */
    double k_risc[3];                 /* T -> S prefactor, 1/s, times min(1, exp(-(S1 - T1)/kT))         */
    double k_rad[3][2];               /* radiative decay [species][spin], 1/s                            */
    double k_nonrad[3][2];            /* non-radiative decay [species][spin], 1/s                        */
    double singlet_fraction;          /* 0.25 (molecule.py:93)                                            */
    double outcoupling;               /* EQE = outcoupling * photons / excitons (reporting only)         */
} hf_x_params;

/* Box totals of the exciton path (u64 each). */
enum {
    HF_XT_GENERATED = 0, HF_XT_EVENTS = 1, HF_XT_FORSTER = 2, HF_XT_DEXTER = 3, HF_XT_ISC = 4, HF_XT_RISC = 5,
    HF_XT_RAD_HOST = 6, HF_XT_RAD_TADF = 7, HF_XT_RAD_FLUO = 8,
    HF_XT_NONRAD_S_HOST = 9, HF_XT_NONRAD_S_TADF = 10, HF_XT_NONRAD_S_FLUO = 11,
    HF_XT_NONRAD_T_HOST = 12, HF_XT_NONRAD_T_TADF = 13, HF_XT_NONRAD_T_FLUO = 14,
    HF_XT_N = 16
};

int hf_x_default_params(hf_x_params *out);
/* Build the neighbour / rate tables, the per-site Boltzmann weights exp(-S1/kT), exp(-T1/kT) of
 * every realisation (needs a lattice) and the exciton state of every trajectory.  Device memory: two f64 arrays of
 * (Z + 2R) X Y sites per realisation, plus — for cutoffs 3, 4, 5 — a copy of both in the padded brick layout the step
 * kernel gathers from ((Z + 2R) (X + 2R) (Y + 2R) sites rounded up to multiples of 4 in x and y; hf_x_launch_info).
 * Calling it again (new parameters, or new energies after hf_upload_lattice) reuses the buffers. */
int hf_x_configure(hf_sim *sim, const hf_x_params *params);
/* Create the first exciton of every trajectory (fills the slot of _form_exciton, reseau.py:434-438). */
int hf_x_spawn(hf_sim *sim);
/* n_events rejection-free KMC events for every trajectory (fills reseau.py:526-550). */
int hf_x_run(hf_sim *sim, int64_t n_events);
/* out[HF_XT_N] summed over realisations; hf_x_read_realisation_tallies: out[n_realisations][HF_XT_N]. */
int hf_x_read_tallies(hf_sim *sim, uint64_t *out);
int hf_x_read_realisation_tallies(hf_sim *sim, uint64_t *out);
/* Tables as the device uses them.  *n_offsets in: capacity, out: M.  offsets[M][3] = dx, dy, dz;
 * g6[M] = r^-6; gd[M] = exp(-2r/L); kf[9], kd[9] = [donor][acceptor] prefactors. */
int hf_x_read_tables(hf_sim *sim, int32_t *n_offsets, int8_t *offsets, double *g6, double *gd, double *kf, double *kd);
/* Per-site Boltzmann weights of one realisation. */
int hf_x_download_weights(hf_sim *sim, int32_t realisation, double *ws1, double *wt1);
/* How hf_x_run is launched (decided by hf_x_configure and, for weights larger than half the L2, measured by the
 * first hf_x_run): trajectories a warp owns at a time, whether the kernel reads the padded 4 x 4 x 1 brick-layout copy
 * of the weights, and the grid.  Results never depend on these.  Any pointer may be NULL. */
int hf_x_launch_info(hf_sim *sim, int32_t *block, int32_t *brick, int32_t *grid);
/* Exciton state of trajectories [first, first + n): site, spin (HF_X_*), time since creation,
 * XTRAJ draws consumed, summed lifetimes of the decayed excitons. */
int hf_x_read_state(hf_sim *sim, int64_t first, int64_t n, int32_t *site, int32_t *spin, double *time,
                    uint64_t *draws, double *lifetime_sum);
/* Trace of the exciton events of the window set with hf_set_trace: kind = HF_EV_FORSTER / HF_EV_MOVE
 * (Dexter) / HF_EV_ISC / HF_EV_DECAY, particule = HF_P_EXCITON, tau = waiting time, spins = channel
 * index, pad[0] = spin of the trajectory's exciton after the step (the new one after a decay). */
int hf_x_read_trace(hf_sim *sim, int64_t trajectory, hf_event *out, int64_t *n);

/* =====================================================================================================
 * High-density exciton devices ("Path XD", BASELINE configs[3]: occupancy blocking and exciton-exciton
 * annihilation).  The reference has NO implementation of this either (parity unpinned): its only
 * occupancy rule is same-type charge blocking (reseau.py:385, 394).  The model is this library's
 * own, specified by oracle/exciton_dense_oracle.c and DESIGN.md section 10.
 * Every trajectory of the handle becomes a DEVICE holding n_excitons interacting excitons on its
 * lattice realisation, with a 2-bit-per-site occupancy mask (0 empty, HF_X_SINGLET, HF_X_TRIPLET).
 * A transfer towards an occupied site is an annihilation channel (rate x annihilation[donor spin
 * index][acceptor spin index]; 0 = pure blocking): the donor is lost, a triplet acceptor hit by a
 * triplet becomes a singlet with probability p_tta; lost excitons are replaced at once on a random
 * empty interior site.  Needs hf_x_configure first (tables, rate constants, Boltzmann weights).
 * ===================================================================================================== */
enum {
    HF_XDT_ANN_SS = 16, HF_XDT_ANN_ST = 17, HF_XDT_ANN_TS = 18, HF_XDT_ANN_TT = 19,   /* donor, acceptor */
    HF_XDT_TTA_UP = 20,               /* triplet acceptors promoted to singlets                          */
    HF_XDT_N = 24                     /* indices 0..14 as HF_XT_*                                         */
};
#define HF_EV_ANNIHILATION 6          /* reported in the `unbound` slot of event.py:49                   */
#define HF_XD_MAX_EXCITONS 1024
int hf_xd_configure(hf_sim *sim, int32_t n_excitons, const double *annihilation /* [2][2] */, double p_tta);
/* Empty the masks and fill every device with n_excitons new excitons. */
int hf_xd_spawn(hf_sim *sim);
/* n_events rejection-free KMC events for every device (one event = one channel of one exciton). */
int hf_xd_run(hf_sim *sim, int64_t n_events);
int hf_xd_read_tallies(hf_sim *sim, uint64_t *out /* [HF_XDT_N] */);
int hf_xd_read_realisation_tallies(hf_sim *sim, uint64_t *out /* [n_realisations][HF_XDT_N] */);
/* One device: site / spin / cached total rate / birth time of its n_excitons slots; scalars[3] =
 * device clock, summed lifetimes of the lost excitons, XDENSE draws consumed.  Any pointer may be NULL. */
int hf_xd_read_device(hf_sim *sim, int64_t device, int32_t *site, int32_t *spin, double *total_rate, double *birth,
                      double *scalars);
/* codes[N]: the occupancy code of every site of one device. */
int hf_xd_read_occupancy(hf_sim *sim, int64_t device, uint8_t *codes);
/* Traces of dense devices are read with hf_x_read_trace: kind as there plus HF_EV_ANNIHILATION,
 * spins = channel | exciton slot << 16, pad[0] = spin of that slot after the event. */

#ifdef __cplusplus
}
#endif
#endif /* HFKMC_H */
