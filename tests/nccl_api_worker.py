"""Worker of tests/test_gpu_parity.py::test_api_nrank_equals_1rank_under_nccl (TEST INFRASTRUCTURE).

Runs the public API calls of the golden tests -- every file-based entry point that shards its work over the ranks -- in
ONE process or under torchrun, and lets rank 0 store what they returned (and what they wrote) in <out>/results.npz.  The
test runs it with 1 and with N ranks on the same fixtures and compares the two files.

    python tests/nccl_api_worker.py <out> <sample_dir> <mini_dir> <water_dir> <slab_dir> <visc_dir> [<c1_dir>]
    python -m torch.distributed.run --nproc-per-node N ... tests/nccl_api_worker.py <same arguments>
"""
import hashlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

MASS = [16.0, 12.01, 1.008, 14.01, 32.06, 16.0, 12.01, 19.0, 24.305]
NUM_MOLS = [591, 66, 33]
NUM_ATOMS = [16, 15, 1]
ELEMENTS = ["O", "C", "H", "N", "S", "O", "C", "F", "Mg"]
REL = [[9, 9, 9, 9], [1, 4, 6, 9]]


def main():
    import torch
    import torch.distributed as dist
    out, sample, mini, water, slab, visc = sys.argv[1:7]
    c1 = sys.argv[7] if len(sys.argv) > 7 else None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")

    def barrier():
        if world > 1:
            dist.barrier()

    def shared_dir(name, copy_from=None):
        d = os.path.join(out, name)
        if rank == 0:
            os.makedirs(d, exist_ok=True)
            if copy_from:
                for f in os.listdir(copy_from):
                    shutil.copy(os.path.join(copy_from, f), d)
        barrier()
        return d

    def files_digest(d, suffix):
        barrier()                                   # every rank has finished writing
        h = hashlib.sha256()
        names = sorted(f for f in os.listdir(d) if f.endswith(suffix))
        for f in names:
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
        return np.array([len(names)]), np.frombuffer(h.digest(), dtype=np.uint8).copy()

    from mdproptools_b200.dynamical.conductivity import Conductivity
    from mdproptools_b200.dynamical.diffusion import Diffusion
    from mdproptools_b200.dynamical.residence_time import ResidenceTime
    from mdproptools_b200.dynamical.viscosity import Viscosity
    import mdproptools_b200.dynamical.viscosity as vmod
    from mdproptools_b200.structural.cluster_analysis import get_clusters
    from mdproptools_b200.structural.hydration_number import get_hydration_number
    from mdproptools_b200.structural.number_density import calc_number_density
    from mdproptools_b200.structural import rdf_cn

    gd = np.load(os.path.join(HERE, "golden", "ref_dynamical.npz"), allow_pickle=False)
    gh = np.load(os.path.join(HERE, "golden", "ref_hydration.npz"), allow_pickle=False)
    mini_mols = gd["mini_num_mols"].tolist()
    R = {}
    pat = os.path.join(sample, "dump.nvt.*.dump")
    # ---- structural: integer counts, host normalisation -> exact --------------------------------------------------
    R["x_atomic_rdf"] = rdf_cn.calc_atomic_rdf(20, 0.05, 9, MASS, REL, pat, save_mode=False).values
    R["x_atomic_rdf_altered"] = rdf_cn.calc_atomic_rdf(12, 0.05, 9, MASS, [[32, 32], [17, 32]], pat, num_mols=NUM_MOLS,
                                                       num_atoms_per_mol=NUM_ATOMS, save_mode=False).values
    R["x_atomic_cn"] = rdf_cn.calc_atomic_cn([2.325, 4.375, 2.375, 13.0], 0.05, 9, MASS, REL, pat, save_mode=False).values
    mrel = [[9, 9, 4], [1, 2, 3]]
    R["x_molecular_rdf"] = rdf_cn.calc_molecular_rdf(20, 0.05, 9, MASS, mrel, pat, NUM_MOLS, NUM_ATOMS, save_mode=False).values
    R["x_molecular_cn"] = rdf_cn.calc_molecular_cn([2.325, 3.775, 4.375], 0.05, 9, MASS, mrel, pat, NUM_MOLS, NUM_ATOMS,
                                                   save_mode=False).values
    R["x_intermolecular_rdf"] = rdf_cn.calc_intermolecular_rdf(20, 0.05, 3, MASS, [[3, 3, 2], [1, 2, 2]],
                                                               os.path.join(mini, "dump.mini.*.dump"), mini_mols, NUM_ATOMS,
                                                               save_mode=False).values
    wd = shared_dir("clusters")
    n = get_clusters(filename=pat, atom_type=32, r_cut=2.3, num_mols=NUM_MOLS, num_atoms_per_mol=NUM_ATOMS, full_trajectory=True,
                     elements=ELEMENTS, alter_atom_types=True, max_force=0.75, working_dir=wd)
    R["x_clusters_count"] = np.array([n])
    R["x_clusters_nfiles"], R["x_clusters_sha"] = files_digest(wd, ".xyz")
    wd = shared_dir("hydration", copy_from=water)
    df = get_hydration_number("dump.water.*.dump", cation_type=1, water_type=2, r_cut=5.0, num_mols=gh["hyd_num_mols"].tolist(),
                              num_atoms_per_mol=gh["hyd_num_atoms"].tolist(), working_dir=wd)
    R["x_hydration_angles"] = df["angles_distribution"].values
    R["x_hydration_factor"] = df["hydration_factor"].values[:1]
    wd = shared_dir("density", copy_from=slab)
    R["x_number_density"] = calc_number_density("dump.slab.*.dump", 1, [2, 3, 1], 0.5, 8.0, "z", working_dir=wd, save_mode=False).values
    wd = shared_dir("residence")
    rt = ResidenceTime([[0, 2.0], [1.9, 2.2], [0, 3.2]], [[32, 32, 1], [1, 27, 1]], os.path.join(mini, "dump.mini.*.dump"), dt=1,
                       num_mols=mini_mols, num_atoms_per_mol=NUM_ATOMS, working_dir=wd)
    rt.calc_auto_correlation()
    R["x_residence"] = rt.corr_df.values
    # ---- dynamical: fp64 sums whose order changes with the rank count -> 1e-12 ------------------------------------
    wd = shared_dir("msd")
    d = Diffusion(timestep=1, units="real", outputs_dir=mini, diff_dir=wd)
    msd, msd_all, msd_int = d.get_msd_from_dump("dump.mini.*.dump", msd_type="allatom", avg_interval=True, tao_coeff=4)
    R["x_msd_all_allatom"], R["t_msd_allatom"], R["t_msd_int_allatom"] = msd_all.values, msd.values, msd_int.values
    msd, msd_all, msd_int = d.get_msd_from_dump("dump.mini.*.dump", msd_type="com", num_mols=mini_mols, num_atoms_per_mol=NUM_ATOMS,
                                                mass=MASS, com_drift=True, avg_interval=True, tao_coeff=4)
    R["t_msd_com"], R["t_msd_all_com"], R["t_msd_int_com"] = msd.values, msd_all.values, msd_int.values
    c = Conductivity("dump.mini.*.dump", mini_mols, NUM_ATOMS, volume=49.182348836183905 ** 3, mass=MASS, temp=298.15, timestep=1,
                     units="real", working_dir=mini)
    j = c.get_charge_flux()
    R["t_cond_flux"] = np.asarray(j)
    R["t_cond_corr"] = np.asarray(c.correlate_charge_flux(j))
    v = Viscosity("log.visc_*", cutoff_time=500, volume=40.0 ** 3, temp=298.15, timestep=1, acf_method="wkt", units="real",
                  working_dir=visc)
    real = vmod.glob.glob
    vmod.glob.glob = lambda p: sorted(real(p))
    try:
        visc_avg, visc_data, acf_data, tvec = v.calc_avg_visc(output_all_data=True)
    finally:
        vmod.glob.glob = real
    R["t_visc_avg"], R["t_visc_acf"] = np.asarray(visc_avg), np.asarray(acf_data)
    if c1:
        c1pat = os.path.join(c1, "dump.nvt.*.dump")
        R["x_c1_rdf"] = rdf_cn.calc_atomic_rdf(20, 0.05, 9, MASS, REL, c1pat, save_mode=False).values
        R["x_c1_cn"] = rdf_cn.calc_atomic_cn([2.325, 4.375, 2.375, 13.0], 0.05, 9, MASS, REL, c1pat, save_mode=False).values
        d = Diffusion(timestep=1, units="real", outputs_dir=c1, diff_dir=wd)
        R["t_c1_msd"] = d.get_msd_from_dump("dump.nvt.*.dump", msd_type="allatom")[0].values
    barrier()
    if rank == 0:
        np.savez(os.path.join(out, "results.npz"), **{k: np.asarray(v_, dtype=np.float64) if np.asarray(v_).dtype.kind == "f" else np.asarray(v_)
                                                     for k, v_ in R.items()})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
