"""Axis cosmetics shared by the plotting helpers (role of mdproptools/utilities/plots.py:13).  Host-only, needs matplotlib."""


def set_axis(ax, axis="both"):
    """Minor ticks halfway between major ones, ``{:g}`` tick labels, inward ticks -- on x, y or both axes."""
    from matplotlib import ticker

    which = {"both": ("x", "y"), "x": ("x",), "y": ("y",)}.get(axis)
    if which is None:
        raise ValueError(f"axis must be 'x', 'y' or 'both', got {axis!r}")
    for name in which:
        a = ax.xaxis if name == "x" else ax.yaxis
        a.set_minor_locator(ticker.AutoMinorLocator(2))
        a.set_major_formatter(ticker.FuncFormatter(lambda v, _: "{:g}".format(v)))
    ax.tick_params(which="major", length=8)
    ax.tick_params(which="minor", length=4)
    ax.tick_params(axis=axis, which="both", direction="in", labelsize=20)
