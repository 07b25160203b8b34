"""Physical constants, same names and values as the reference's ``constants.py:3-6``.

The expressions are kept (not the rounded literals): ``8.617333262 * 10.**(-5)`` is one ulp
above ``8.617333262e-5``, and the device code embeds exactly these doubles
(``csrc/hf_frm.cuh``: ``HF_KT``, ``HF_ELECTROSTATIC``).
"""
from math import pi

BOLTZMANN: float = 8.617333262 * 10.**(-5)                                            # [eV/K]
VACUUM_PERMITTIVITY: float = 55263.49406                                             # [e^2/(eV nm)]
RELATIVE_PERMITTIVITY: float = 3.
ELECTROSTATIC: float = 1. / (4. * pi * VACUUM_PERMITTIVITY * RELATIVE_PERMITTIVITY)  # [eV nm]

assert BOLTZMANN.hex() == "0x1.696fe94e12d2dp-14" and ELECTROSTATIC.hex() == "0x1.01b112ab38751p-21"
assert (BOLTZMANN * 300.).hex() == "0x1.a78f25677e0f1p-6"
