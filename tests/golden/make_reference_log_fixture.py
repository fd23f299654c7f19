"""Generates tests/golden/reference_run_logs.json from the run logs the reference ships next to its binaries
(OpenMP/NGP/src/output.txt, OpenMP/CIC/src/output.txt): what the REAL reference program printed at start-up for its default
He RF decks -- time step / plasma frequency, grid, particle masses and weights, and the null-collision sets it built from
PIC1D/CollisionData (reduced masses, merged table length, max sigma*v, collision order).  These are the only outputs of the
Fortran binary available here (it cannot be built: no Fortran compiler, Intel-only modules), so they are the golden vectors
the host-side logic and the MCC constructor are pinned on (tests/test_reference_log_pins.py).
Run in the build container only (reads /root/reference):  python tests/golden/make_reference_log_fixture.py"""
import json
import os
import re

REF = "/root/reference/OpenMP"
HERE = os.path.dirname(os.path.abspath(__file__))
NUM = r"([-+]?\d+\.?\d*(?:[EeDd][-+]?\d+)?)"


def first(txt, label, cast=float):
    m = re.search(re.escape(label) + r"\s*" + NUM, txt)
    if not m:
        raise KeyError(label)
    return cast(m.group(1).replace("D", "E").replace("d", "e"))


def parse(path):
    txt = open(path).read()
    out = dict(source=os.path.relpath(path, "/root/reference"))
    out["numThread"] = first(txt, "Number of threads is:", int)
    out["n_ave"] = first(txt, "Average initial electron density:")
    out["T_e"] = first(txt, "Initial electron temperature:")
    out["T_i"] = first(txt, "Initial ion temperature:")
    out["fraction_of_inverse_plasma_frequency"] = first(txt, "Fraction of 1/w_p for time step:")
    out["del_t"] = first(txt, "del_t will be:")
    out["NumberXNodes"] = first(txt, "Number of nodes:", int)
    out["NumberXHalfNodes"] = first(txt, "Number of half nodes:", int)
    out["L_domain"] = first(txt, "Grid length:")
    out["gridType"] = first(txt, "gridType:", int)
    out["bcL"], out["bcR"] = first(txt, "Left boundary type:", int), first(txt, "Right boundary type:", int)
    out["smallest_delX"] = first(txt, "smallest delX:")
    out["RF_frequency"] = first(txt, "RF frequency:")
    out["RF_half_amplitude"] = first(txt, "RF_half_amplitude:")
    out["RF_rad_frequency"] = first(txt, "RF_rad_frequency:")
    parts = []
    for m in re.finditer(r"Initializing (\S+)\s+Amount of macroparticles is:\s*(\d+)\s+Particle mass is:\s*" + NUM + r"\s+Particle charge is:\s*" + NUM
                         + r"\s+Particle weight is:\s*" + NUM + r"\s+Particle mean KE is:\s*" + NUM, txt):
        parts.append(dict(name=m.group(1), macroparticles=int(m.group(2)), mass=float(m.group(3)), charge=float(m.group(4)), weight=float(m.group(5)),
                          mean_KE_eV=float(m.group(6))))
    out["particles"] = parts
    m = re.search(r"Initializing target particle:(\S+)\s+Particle mass is:\s*" + NUM + r"\s+Particle temperature\(K\) is:\s*" + NUM
                  + r"\s+Particle density is:\s*" + NUM, txt)
    out["target"] = dict(name=m.group(1), mass=float(m.group(2)), temperature=float(m.group(3)), density=float(m.group(4)))
    sets = []
    for blk in txt.split("Null Collision generated")[1:]:
        blk = blk.split("---------------------------")[0]
        s = dict(reactants=re.search(r"Reactants are:\s*(\S+) and (\S+)", blk).groups())
        s["numberCollisions"] = first(blk, "Amount of collisions:", int)
        s["reducedMass"] = first(blk, "Reduced mass:")
        s["lengthArrays"] = first(blk, "length of arrays:", int)
        s["sigmaVMax"] = first(blk, "Max sigma_v:")
        s["reducedMassIonization"] = first(blk, "reducedMassIonization:")
        s["collisions"] = [dict(energyThreshold=float(a), collisionType=int(b)) for a, b in
                           re.findall(r"Energy threshold:\s*" + NUM + r"\s*\n\s*Collision type:\s*(\d+)", blk)]
        sets.append(s)
    out["nullCollisions"] = sets
    m = re.search(r"Absolute error:\s*" + NUM + r"\s*\n\s*Relative error:\s*" + NUM, txt)
    if m:                                                  # older logs print the solver inputs differently
        out["eps_a"], out["eps_r"] = float(m.group(1)), float(m.group(2))
        out["m_Anderson"] = first(txt, "Anderson number m is:", int)
        out["Beta_k"] = first(txt, "Relaxation parameter is:")
    # the diagnostic dumps: particle counts as the plasma relaxes (statistical: the reference seeds from the clock)
    out["electrons_at_dumps"] = [int(x) for x in re.findall(r"Number of electrons is:\s*(\d+)", txt)]
    out["ions_at_dumps"] = [int(x) for x in re.findall(r"Number of ions is:\s*(\d+)", txt)]
    return out


if __name__ == "__main__":
    # the third log (left in the NGP_Periodic source directory) is another bounded NGP run, on 2001 nodes
    data = dict(ngp=parse(os.path.join(REF, "NGP/src/output.txt")), cic=parse(os.path.join(REF, "CIC/src/output.txt")),
                ngp_2001=parse(os.path.join(REF, "NGP_Periodic/src/output.txt")))
    with open(os.path.join(HERE, "reference_run_logs.json"), "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(data["ngp"], indent=1)[:1500])
