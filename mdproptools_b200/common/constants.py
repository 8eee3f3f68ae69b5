"""Physical constants and LAMMPS-unit -> SI factors.

The VALUES are part of the numerical contract with the reference (mdproptools/common/constants.py:22-168):
every expression below evaluates, in IEEE double arithmetic, to the same bits as the reference's tables
(same literals, same operation order), because MSD / conductivity / viscosity results are multiplied by them.
"""

BOLTZMANN = 1.380649 * 10 ** -23  # J/K            (constants.py:22)
ELEMENTARY_CHARGE = 1.602176634 * 10 ** -19  # C   (constants.py:24)
AVOGADRO = 6.02214076 * 10 ** 23  # 1/mol          (constants.py:26)
LIGHT_SPEED = 299792458  # m/s
BOHR_RADIUS = 5.29177210903 * 10 ** -11  # m
CAL_TO_J = 4.184
HA_TO_J = 4.3597447222071 * 10 ** -18

SUPPORTED_UNITS = ["real", "metal", "si", "cgs", "electron", "micro", "nano"]


def _table(real, metal, si, cgs, electron, micro, nano):
    return {"real": real, "metal": metal, "si": si, "cgs": cgs, "electron": electron, "micro": micro, "nano": nano}


MASS_CONVERSION = _table(10 ** -3 / AVOGADRO, 10 ** -3 / AVOGADRO, 1, 10 ** -3, 10 ** -3 / AVOGADRO,
                         10 ** -3 * 10 ** -12, 10 ** -3 * 10 ** -18)
DISTANCE_CONVERSION = _table(10 ** -10, 10 ** -10, 1, 10 ** -2, BOHR_RADIUS, 10 ** -6, 10 ** -9)
TIME_CONVERSION = _table(10 ** -15, 10 ** -12, 1, 1, 10 ** -15, 10 ** -6, 10 ** -9)
ENERGY_CONVERSION = _table(10 ** 3 * CAL_TO_J / AVOGADRO, ELEMENTARY_CHARGE, 1, 10 ** -7, HA_TO_J,
                           MASS_CONVERSION["micro"], MASS_CONVERSION["nano"])
VELOCITY_CONVERSION = _table(
    DISTANCE_CONVERSION["real"] / TIME_CONVERSION["real"],
    DISTANCE_CONVERSION["metal"] / TIME_CONVERSION["metal"],
    1,
    DISTANCE_CONVERSION["cgs"] / TIME_CONVERSION["cgs"],
    DISTANCE_CONVERSION["electron"] / (1.03275 * 10 ** -15),
    DISTANCE_CONVERSION["micro"] / TIME_CONVERSION["micro"],
    DISTANCE_CONVERSION["nano"] / TIME_CONVERSION["nano"],
)
FORCE_CONVERSION = _table(*[
    1 if u == "si" else ENERGY_CONVERSION[u] / DISTANCE_CONVERSION[u] for u in SUPPORTED_UNITS
])
TORQUE_CONVERSION = ENERGY_CONVERSION
TEMPERATURE_CONVERSION = _table(1, 1, 1, 1, 1, 1, 1)
PRESSURE_CONVERSION = _table(101325, 10 ** 5, 1, 10 ** -6 * 10 ** 5, 1,
                             ENERGY_CONVERSION["micro"] / DISTANCE_CONVERSION["micro"] ** 3,
                             ENERGY_CONVERSION["nano"] / DISTANCE_CONVERSION["nano"] ** 3)
VISCOSITY_CONVERSION = _table(0.1, 0.1, 1, 0.1, 1,
                              PRESSURE_CONVERSION["micro"] * TIME_CONVERSION["micro"],
                              PRESSURE_CONVERSION["nano"] * TIME_CONVERSION["nano"])
CHARGE_CONVERSION = _table(ELEMENTARY_CHARGE, ELEMENTARY_CHARGE, 1, 1 / 10 / LIGHT_SPEED, ELEMENTARY_CHARGE,
                           10 ** -12, ELEMENTARY_CHARGE)
DIPOLE_CONVERSION = _table(
    CHARGE_CONVERSION["real"] * DISTANCE_CONVERSION["real"],
    CHARGE_CONVERSION["metal"] * DISTANCE_CONVERSION["metal"],
    1,
    CHARGE_CONVERSION["cgs"] * DISTANCE_CONVERSION["cgs"],
    10 ** -21 / LIGHT_SPEED,
    CHARGE_CONVERSION["micro"] * DISTANCE_CONVERSION["micro"],
    CHARGE_CONVERSION["nano"] * DISTANCE_CONVERSION["nano"],
)
ELECTRIC_FIELD_CONVERSION = _table(
    1 / DISTANCE_CONVERSION["real"], 1 / DISTANCE_CONVERSION["metal"], 1,
    FORCE_CONVERSION["cgs"] / CHARGE_CONVERSION["cgs"], 1 / 10 ** -2,
    1 / DISTANCE_CONVERSION["micro"], 1 / DISTANCE_CONVERSION["nano"],
)
DENSITY_3D_CONVERSION = {
    u: (1 if u == "si" else
        MASS_CONVERSION["cgs"] / DISTANCE_CONVERSION["cgs"] ** 3 if u in ("real", "metal", "cgs") else
        MASS_CONVERSION[u] / DISTANCE_CONVERSION[u] ** 3)
    for u in ("real", "metal", "si", "cgs", "micro", "nano")
}
