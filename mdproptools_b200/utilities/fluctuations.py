"""Fluctuation of a thermo-log property over time (role of mdproptools/utilities/fluctuations.py:14-57).

Returns (mean, sample standard deviation) of ``log[log_prop]`` -- pandas ``describe()`` semantics, ddof = 1 -- prints
them, and saves the time trace with its mean line.  Host-only (a thermo log is a few MB); matplotlib is needed for the
figure only and is imported lazily.
"""
from __future__ import annotations

import os

import numpy as np

from ..common import constants


def _get_stats(stats):
    return "(" + ", ".join([f"{k}:{v: .4g}" for k, v in stats.items()]) + ")"


def fluctuation_stats(log, log_prop):
    v = np.asarray(log[log_prop], dtype=np.float64)
    return {"mean": float(v.mean()), "std": float(v.std(ddof=1)) if len(v) > 1 else float("nan")}


def plot_fluctuations(log, log_prop, title, filename, timestep=1, units="real", working_dir=None):
    working_dir = working_dir or os.getcwd()
    stats = fluctuation_stats(log, log_prop)
    print("{}: mean = {}, std = {}".format(log_prop, stats["mean"], stats["std"]))
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError as exc:
        raise ImportError("plot_fluctuations needs matplotlib for the figure; use fluctuation_stats for the numbers") from exc
    from .plots import set_axis

    y = np.asarray(log[log_prop], dtype=np.float64)
    t = np.asarray(log["Step"], dtype=np.float64) * timestep * constants.TIME_CONVERSION[units] * 10**9
    fig, ax = plt.subplots(figsize=(8, 6))
    set_axis(ax, axis="both")
    ax.plot(t, y, linewidth=2, color="red")
    ax.axhline(stats["mean"], linewidth=2, color="#000000", ls="--")
    ax.set_title("{} {}".format(title, _get_stats(stats)), fontsize=18)
    ax.set_xlabel(r"$\mathrm{Time, 10^9 (m^2/s)}$", fontsize=18)       # the reference's label (fluctuations.py:48)
    ax.set_xlim(0, None)
    lo, hi = float(y.min()), float(y.max())
    ax.set_ylim(lo * 2 if lo < 0 else lo / 2, hi * 2 if hi > 0 else -hi * 2)
    fig.tight_layout(pad=3)
    fig.savefig(f"{working_dir}/{filename}", bbox_inches="tight", pad_inches=0.1)
    plt.close(fig)
    return stats["mean"], stats["std"]
