"""Green-Kubo shear viscosity from the off-diagonal pressure tensor -- drop-in for
``mdproptools.dynamical.viscosity.Viscosity`` (reference mdproptools/dynamical/viscosity.py; citations are
lines of that file).

Device work (csrc/corr.cu): the pressure-tensor autocorrelations of all replicates and all three components
in one batched direct correlation (autocorrelate, :86-120) and their running trapezoid integrals scaled by
V / (k_B T) (calc_visc, :139-153).  The double-exponential fit and the bootstrapping (:239-434) are scipy
curve fits on a few thousand points and stay on the host.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from .. import dist, ops
from ..common import constants
from ..io.log import parse_lammps_log

TENSOR_LABELS = ["Pxy", "Pxz", "Pyz"]  # :27


class Viscosity:
    def __init__(self, log_pattern, cutoff_time, volume, temp=298.15, timestep=1, acf_method="wkt", units="real",
                 working_dir=None):
        self.log_pattern = log_pattern
        self.cutoff_time = cutoff_time
        self.units = units
        self.volume = volume * constants.DISTANCE_CONVERSION[self.units] ** 3
        self.temp = temp
        self.timestep = timestep
        self.acf_method = acf_method
        self.working_dir = working_dir or os.getcwd()
        self.time = None
        self.step_to_s = self.timestep * constants.TIME_CONVERSION[self.units]

    @staticmethod
    def autocorrelate(series, method):
        """Unbiased time autocorrelation (:86-120).  Both of the reference's methods ("wkt" FFT and
        "brute_force" np.correlate) define the same quantity; here it is always the direct fp64 sum on the
        device."""
        if method not in ("brute_force", "wkt"):
            raise ValueError("Method string input not recognized")
        s = torch.as_tensor(np.ascontiguousarray(series, dtype=np.float64)).cuda().reshape(1, -1)
        return ops.xcorr_unbiased(s, s)[0].cpu().numpy()

    @staticmethod
    def exp_func(t, A, alpha, tau1, tau2):
        """Double exponential running-integral model (:122-137)."""
        return A * alpha * tau1 * (1 - np.exp(-t / tau1)) + A * (1 - alpha) * tau2 * (1 - np.exp(-t / tau2))

    def calc_visc(self, acf, dt):
        """(:139-153) cumulative trapezoid (no leading zero) times V / (k_B T)."""
        y = torch.as_tensor(np.ascontiguousarray(acf, dtype=np.float64)).cuda().reshape(1, -1)
        integral = ops.cumtrapz(y, dt, 1.0, leading_zero=False)[0].cpu().numpy()
        return np.multiply(self.volume / (constants.BOLTZMANN * self.temp), integral)

    def _calc_3d_visc_batched(self, log_dfs):
        """All replicates x 3 tensor components in one correlation call and one integration call."""
        if self.units not in constants.SUPPORTED_UNITS:
            raise KeyError("Unit type not supported. Supported units are: " + str(constants.SUPPORTED_UNITS))
        T = len(log_dfs[0])
        if any(len(df) != T for df in log_dfs):
            return None
        series = np.stack([df[label].to_numpy(dtype=np.float64) for df in log_dfs for label in TENSOR_LABELS])
        # replicates (channels) are split over ranks; the rows are gathered with one all-reduce
        C = series.shape[0]
        lo, hi = dist.shard_range(C)
        dev = torch.device("cuda")
        acf = torch.zeros((C, T), dtype=torch.float64, device=dev)
        if hi > lo:
            s = torch.from_numpy(np.ascontiguousarray(series[lo:hi])).to(dev)
            acf[lo:hi] = ops.xcorr_unbiased(s, s) * constants.PRESSURE_CONVERSION[self.units] ** 2   # (:181-183)
        dist.all_reduce_sum_(acf)
        out = []
        for r, df in enumerate(log_dfs):
            time_data = df["Step"] * self.step_to_s
            delta_t = time_data.iloc[1] - time_data.iloc[0]
            a = acf[3 * r:3 * r + 3].contiguous()
            integral = ops.cumtrapz(a, float(delta_t), 1.0, leading_zero=False).cpu().numpy()
            viscosity_data = np.multiply(self.volume / (constants.BOLTZMANN * self.temp), integral)
            out.append((np.mean(viscosity_data, axis=0), viscosity_data, a.cpu().numpy()))
        return out

    def _calc_3d_visc(self, log_df):
        """(:155-191) one replicate."""
        return self._calc_3d_visc_batched([log_df])[0]

    def calc_avg_visc(self, output_all_data=False):
        """(:193-237) parse the replicate logs, drop everything before ``cutoff_time``, return the running
        viscosity of every replicate."""
        list_log_df = []
        log_files = glob.glob(f"{self.working_dir}/{self.log_pattern}")
        for file in log_files:
            list_log_df.append(parse_lammps_log(file)[0])
        first = list_log_df[0]
        cutoff_time_idx = first.index.get_loc(first[first["Step"] == self.cutoff_time].index[0])
        cut = [df.iloc[cutoff_time_idx:] for df in list_log_df]
        res = self._calc_3d_visc_batched(cut)
        if res is None:   # replicates of different length: one call each
            res = [self._calc_3d_visc(df) for df in cut]
        visc_avg = [r[0] for r in res]
        visc_data = [r[1] for r in res]
        acf_data = [r[2] for r in res]
        self.time = np.array(list_log_df[0]["Step"][: len(visc_avg[0]) - 1]) * self.timestep
        if output_all_data:
            return visc_avg, visc_data, acf_data, self.time
        return visc_avg

    def fit_avg_visc(self, visc_avg, initial_guess=[1e-10, 0.8, 1.1e4, 1.1e4], plot=False, plot_file="viscosity.png"):
        """(:239-380) average/std over replicates, weighted double-exponential fit between 2 ps and the time where
        std >= 0.4 * viscosity; returns the infinite-time viscosity A alpha tau1 + A (1-alpha) tau2."""
        from scipy import optimize

        visc = np.average(visc_avg, axis=0)
        std = np.std(visc_avg, axis=0)
        time_indexes = np.where(self.time > 2000)
        idx_start_time = time_indexes[0][0] if time_indexes else 1
        std_indexes = np.where(std >= 0.4 * visc)
        idx_cut_time = std_indexes[0][0] if std_indexes else 1
        sl = slice(idx_start_time, idx_cut_time)
        popt2, pcov2 = optimize.curve_fit(
            self.exp_func, self.time[sl], visc[sl], sigma=1 / std[sl] ** 0.5,
            bounds=(0, [max(visc[sl]), 1, 5 * self.time[idx_cut_time], 5 * self.time[idx_cut_time]]),
            p0=initial_guess, maxfev=1000000,
        )
        viscosity = popt2[0] * popt2[1] * popt2[2] + popt2[0] * (1 - popt2[1]) * popt2[3]
        if plot:
            try:
                import matplotlib
                matplotlib.use("Agg")
                import matplotlib.pyplot as plt
            except ImportError as exc:
                raise ImportError("plot=True needs matplotlib") from exc
            fig, ax = plt.subplots(1, 3, figsize=(18, 5))
            for v in visc_avg:
                ax[0].plot(self.time, v[: len(self.time)], linewidth=1)
            ax[0].plot(self.time, visc[: len(self.time)], color="black", linewidth=2)
            ax[0].axvline(self.time[idx_cut_time], color="black", linestyle="--")
            ax[1].plot(self.time, std[: len(self.time)])
            ax[2].plot(self.time[sl], visc[sl])
            ax[2].plot(self.time[sl], self.exp_func(self.time[sl], *popt2), color="black", linestyle="--")
            fig.savefig(f"{self.working_dir}/{plot_file}", bbox_inches="tight", pad_inches=0.1)
            plt.close(fig)
        return viscosity

    def bootstrapping(self, visc_avg, num_replicates, tot_replicates, initial_guess=[1e-10, 0.8, 1.1e4, 1.1e4], plot=True,
                      seed=None):
        """(:382-434) ``tot_replicates`` bootstrap iterations, each refitting the mean of ``num_replicates`` running integrals
        drawn WITHOUT replacement (``random.sample``, as the reference: no replicate twice within one iteration); returns
        ``(mean viscosity, std)`` of the fitted values.  ``seed`` (an addition, trailing keyword) makes the draw reproducible;
        None uses the global ``random`` state exactly as the reference does."""
        import random
        rnd = random.Random(seed) if seed is not None else random
        idx = np.zeros((tot_replicates, num_replicates), dtype=int)
        for i in range(tot_replicates):
            idx[i] = rnd.sample(range(len(visc_avg)), num_replicates)
        visc_samples = np.array(visc_avg)[idx]
        all_visc = []
        for ind, visc in enumerate(visc_samples):
            print(f"Fitting viscosity sample {ind + 1} out of {len(visc_samples)}")
            all_visc.append(self.fit_avg_visc(visc_avg=visc, initial_guess=initial_guess, plot=plot,
                                              plot_file=f"viscosity_{ind + 1}.png"))
        return np.average(all_visc), np.std(all_visc)
