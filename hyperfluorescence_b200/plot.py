"""3-D scatter of the carriers, same call signature as the reference's ``plot.plot``
(``plot.py:6-138``): electrons blue, holes red, excitons magenta, marker by molecule class,
saved as ``<name>.png``.  Pure visualisation, not on any compute path; matplotlib is imported
lazily so that ``from plot import plot`` (``OLED.py:3``) works where it is not installed.
"""
from __future__ import annotations

from .molecule import Fluorescent, Host, TADF

_MARKERS = {Host: "o", TADF: "^", Fluorescent: "*"}
_COLOURS = ("blue", "red", "magenta")


def plot(electrons, holes, excitons, x_max: int, y_max: int, z_max: int, name: str = "lattice") -> None:
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError("plot() needs matplotlib, which is not installed") from exc
    fig = plt.figure()
    ax = fig.add_subplot(projection="3d")
    for group, colour in zip((electrons, holes, excitons), _COLOURS):
        for cls, marker in _MARKERS.items():
            pts = [p for p, kind in group if kind is cls]
            if pts:
                ax.scatter([p.x for p in pts], [p.y for p in pts], [p.z for p in pts], c=colour, marker=marker)
    ax.set_xlim(0, x_max)
    ax.set_ylim(0, y_max)
    ax.set_zlim(0, z_max)
    fig.savefig(f"{name}.png")
    plt.close(fig)
