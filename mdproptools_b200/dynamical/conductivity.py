"""Green-Kubo ionic conductivity -- drop-in for ``mdproptools.dynamical.conductivity.Conductivity``
(reference mdproptools/dynamical/conductivity.py and _conductivity.py; citations are lines of
conductivity.py unless prefixed).

Device work: the charge-flux stage (per frame: molecular COM velocity, molecular charge, per-type
J = sum q v, _conductivity.py:7-36) as a streaming segmented reduction over ``[T][3][N]`` velocities
(csrc/reduce.cu), the ntypes^2 x 3 unbiased cross-correlations (:97-114, :197-214) as batched direct fp64
correlations and the cumulative trapezoid (:216-232) (csrc/corr.cu).  ``detect_time_range`` (:116-165) is
discrete window selection and stays on the host, fed with the device results.

Multi-GPU: frames are split round-robin for the flux stage (all-reduce of the zero-padded J), the
correlation channels are split over ranks and summed with one fp64 all-reduce.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from .. import dist, ops
from ..common import constants
from ..common.com_mols import atom_masses, mol_membership
from ..io import dump as _dump
from ..io.pipeline import FrameBatches


class Conductivity:
    def __init__(self, filename, num_mols, num_atoms_per_mol, volume, mass=None, temp=298.15, timestep=1, units="real",
                 working_dir=None):
        self.working_dir = working_dir or os.getcwd()
        self.filename = filename
        self.mass = mass
        self.num_mols = num_mols
        self.num_atoms_per_mol = num_atoms_per_mol
        self.units = units
        self.volume = volume * constants.DISTANCE_CONVERSION[self.units] ** 3  # volume in m^3 (:92)
        self.temp = temp
        self.timestep = timestep
        self.time = []  # time data used to calculate GK integral

    # -- correlators --------------------------------------------------------------------------------
    @staticmethod
    def correlate(a, b):
        """Unbiased correlation c[tau] = sum_t a[t+tau] b[t] / (T - tau) (:97-114), direct sum on the device."""
        a = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda().reshape(1, -1)
        b = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).cuda().reshape(1, -1)
        return ops.xcorr_unbiased(a, b)[0].cpu().numpy()

    @staticmethod
    def detect_time_range(flux, tol):
        """Host restatement of :116-165: block-wise std of the correlation function, normalised by the std of
        the block stds, thresholded at ``tol``, median-filtered; the longest run of quiet blocks wins."""
        flux = pd.Series(flux, name="flux")
        time_step = max(int(len(flux) / 10000), 5)
        ind = [i // time_step for i in range(len(flux))]
        flux_std = flux.groupby(ind).transform("std")
        std = flux_std.std()
        div = std if std else 1
        flux_std = flux_std / div
        flux_std = (flux_std < tol).astype("int").to_frame()
        flux_std = (
            flux_std.rolling(window=4 * time_step + 1, min_periods=3 * time_step + 1, center=True)
            .median().fillna(0)["flux"].to_list()
        )
        s_e_list = []
        found_start = False
        for k, v in enumerate(flux_std):
            if v == 1 and not found_start:
                s_e_list.append((k,))
                found_start = True
            elif v < 1 and found_start:
                s_e_list[-1] = s_e_list[-1] + (k,)
                found_start = False
        if s_e_list and len(s_e_list[-1]) == 1:
            s_e_list[-1] = s_e_list[-1] + (len(flux_std) - 1,)
        max_s_e = 0
        max_s_e_ind = None
        for s_e_ind, s_e in enumerate(s_e_list):
            if s_e[1] - s_e[0] > max_s_e:
                max_s_e = s_e[1] - s_e[0]
                max_s_e_ind = s_e_ind
        return s_e_list[max_s_e_ind]   # TypeError when no window exists, as in the reference

    # -- pipeline stages ------------------------------------------------------------------------------
    def get_charge_flux(self):
        """J[3, ntypes, T] (:167-195)."""
        pattern = f"{self.working_dir}/{self.filename}"
        cols = _dump.available_columns(pattern)
        want = ["id", "type", "q", "vx", "vy", "vz"]
        if not self.mass:
            assert "mass" in cols, "Missing atom masses in dump file."
            want.append("mass")
        mol_type, _, seg_off = mol_membership(self.num_mols, self.num_atoms_per_mol)
        ntypes = len(self.num_mols)
        type_off = np.concatenate(([0], np.cumsum(self.num_mols))).astype(np.int64)
        w, r = dist.world_size(), dist.rank()
        sel = (lambda i: i % w == r) if w > 1 else None
        batches = FrameBatches(pattern, want, frame_select=sel)
        pieces, times = [], {}
        dev = torch.device("cuda", torch.cuda.current_device())    # a rank without frames still joins the collectives
        seg_d = torch.from_numpy(seg_off.astype(np.int32)).to(dev)
        vs, qs = constants.VELOCITY_CONVERSION[self.units], constants.CHARGE_CONVERSION[self.units]
        for batch in batches:
            d = batch.wait()
            host = batch.host.numpy()
            n = host.shape[2]
            if seg_off[-1] != n:
                raise ValueError(f"Length of values ({seg_off[-1]}) does not match length of index ({n})")
            mcol = None if self.mass else batch.col("mass")

            def masses(k):
                return atom_masses(host[k, 1], self.mass) if self.mass else host[k, mcol]

            vel = d[:, 3:6, :].contiguous()
            # the kernel sums each molecule's charge and mass once per call; the reference does it per frame
            # (_conductivity.py:11-18), so one call serves the batch only when q, type and mass do not change inside it
            # (fixed-charge force fields); fluctuating charges (QEq, polarisable models) take one call per frame
            static = bool((host[:, 2] == host[0, 2]).all() and (host[:, 1] == host[0, 1]).all()
                          and (mcol is None or (host[:, mcol] == host[0, mcol]).all()))
            if static:
                j = ops.charge_flux(vel, torch.from_numpy(np.ascontiguousarray(masses(0))).to(dev), d[0, 2].contiguous(),
                                    seg_d, type_off, vs, qs)
            else:
                j = torch.cat([ops.charge_flux(vel[k:k + 1], torch.from_numpy(np.ascontiguousarray(masses(k))).to(dev),
                                               d[k, 2].contiguous(), seg_d, type_off, vs, qs)
                               for k in range(vel.shape[0])], dim=2)
            pieces.append(([m.index for m in batch.metas], j))
            for m in batch.metas:
                times[m.index] = m.timestep * constants.TIME_CONVERSION[self.units] * self.timestep
        T = batches.total_frames or 0
        if T == 0:
            raise ValueError(f"no dump frames found for {pattern!r}")
        J = torch.zeros((3, ntypes, T), dtype=torch.float64, device=dev)
        tt = torch.zeros((T,), dtype=torch.float64, device=dev)
        for idxs, j in pieces:
            ii = torch.tensor(idxs, device=dev)
            J[:, :, ii] = j
            tt[ii] = torch.tensor([times[i] for i in idxs], dtype=torch.float64, device=dev)
        dist.all_reduce_sum_(J)
        dist.all_reduce_sum_(tt)
        self.time = tt.cpu().tolist()
        self._flux_dev = J
        return J.cpu().numpy()

    def correlate_charge_flux(self, flux):
        """tot_flux[i] = sum_j sum_c corr(J_c,i , J_c,j); last row = grand total (:197-214)."""
        flux_d = torch.as_tensor(np.ascontiguousarray(flux, dtype=np.float64)).cuda()
        _, ntypes, T = flux_d.shape
        chans = [(k, i, j) for i in range(ntypes) for j in range(ntypes) for k in range(3)]
        lo, hi = dist.shard_range(len(chans))
        tot = torch.zeros((ntypes + 1, T), dtype=torch.float64, device=flux_d.device)
        if hi > lo:
            a = torch.stack([flux_d[k, i] for (k, i, j) in chans[lo:hi]]).contiguous()
            b = torch.stack([flux_d[k, j] for (k, i, j) in chans[lo:hi]]).contiguous()
            corr = ops.xcorr_unbiased(a, b)
            for row, (k, i, j) in zip(corr, chans[lo:hi]):       # fixed order: i, then j, then axis
                tot[i] += row
                tot[-1] += row
        dist.all_reduce_sum_(tot)
        return tot.cpu().numpy()

    def integrate_charge_flux_correlation(self, tot_flux):
        """Cumulative trapezoid with a leading zero, dx = time[1] - time[0] (:216-232)."""
        delta = self.time[1] - self.time[0]
        y = torch.as_tensor(np.ascontiguousarray(tot_flux, dtype=np.float64)).cuda()
        return ops.cumtrapz(y, delta, 1.0, leading_zero=True).cpu().numpy()

    def fit_curve(self, tot_flux, integral, tol):
        """(:234-257)"""
        ave = np.zeros((len(integral)))
        time_range = np.zeros((len(integral)), dtype=object)
        for i in range(len(integral)):
            time_range_ind = self.detect_time_range(tot_flux[i], tol=tol)
            ave[i] = np.average(integral[i][time_range_ind[0]: time_range_ind[1]])
            time_range[i] = (self.time[time_range_ind[0]], self.time[time_range_ind[1]])
        return ave, time_range

    def green_kubo(self, ave):
        """sigma = <integral> / 3 / k_B / T / V (:259-274)."""
        cond = np.zeros((len(ave)))
        for i in range(len(ave)):
            cond[i] = ave[i] / 3 / constants.BOLTZMANN / self.temp / self.volume
        return cond

    def calc_cond(self, tol=1e-4, plot=False, save=False):
        """Wrapper (:276-398): flux -> correlation -> integral -> plateau window -> conductivity (S/m)."""
        j = self.get_charge_flux()
        tot_flux = self.correlate_charge_flux(j)
        integral = self.integrate_charge_flux_correlation(tot_flux)
        ave, time_range = self.fit_curve(tot_flux, integral, tol)
        cond = self.green_kubo(ave)
        if plot:
            self._plot(tot_flux, integral, time_range)
        if save and dist.rank() == 0:
            charge_flux = np.append(np.array([self.time]), tot_flux, axis=0)
            integral_t = np.append(np.array([self.time]), integral, axis=0)
            start_time = [i[0] for i in time_range]
            end_time = [i[1] for i in time_range]
            cond_t = np.asarray([start_time, end_time, cond])
            mol_names = ",".join([str(i + 1) for i in range(len(tot_flux) - 1)]) + ",tot"
            col_names = "t" + "," + mol_names
            np.savetxt(f"{self.working_dir}/charge_flux.csv", charge_flux.T, delimiter=",", header=col_names, comments="")
            np.savetxt(f"{self.working_dir}/integral.csv", integral_t.T, delimiter=",", header=col_names, comments="")
            np.savetxt(f"{self.working_dir}/conductivity.csv", cond_t.T, delimiter=",", header="start_t,end_t,cond",
                       comments="")
            cond = cond_t   # the reference returns the stacked array when save=True (:376)
        return cond

    def _plot(self, tot_flux, integral, time_range):
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError as exc:
            raise ImportError("plot=True needs matplotlib") from exc
        t = np.array(self.time) * 10 ** 9
        fig, ax = plt.subplots(1, 2, figsize=(20, 5))
        for i in range(len(tot_flux) - 1):
            ax[0].plot(t, tot_flux[i], linewidth=2)
            ax[1].plot(t, integral[i], linewidth=2, label=i + 1)
        ax[0].plot(t, tot_flux[-1], linewidth=2, color="black")
        ax[1].plot(t, integral[-1], linewidth=2, color="black", label="total")
        for a in ax:
            a.axvline(time_range[-1][0] * 10 ** 9, color="black", linestyle="--")
            a.axvline(time_range[-1][1] * 10 ** 9, color="black", linestyle="--")
            a.set_xscale("log")
            a.set_xlabel("Time, 10^9 (s)")
        ax[1].legend(frameon=False)
        fig.savefig(f"{self.working_dir}/conductivity.png", bbox_inches="tight", pad_inches=0.1)
        plt.close(fig)

    def einstein(self):
        pass

    def nernst(self):
        pass
