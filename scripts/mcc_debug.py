import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gpu_mcc import *
from implicitpic1d_b200 import _lib
o, (e, ion), seeds, ncs_o, solver, plist, tl, ncs_g = ar_states()
del_t = 1e-10
for it in range(20):
    for s in (e, ion):
        mp.set_num_to_collide(o, s)
    for r in ncs_o:
        r.generateCollision(seeds, del_t)
    for p in plist:
        _lib.check(_lib.load().ipic_set_num_to_collide(solver._h, p._index, p.count()))
    for g in ncs_g:
        g.generateCollision(plist, tl, del_t=del_t)
    for g, r in zip(ncs_g, ncs_o):
        if it % 5 != 4: continue
        to = r.totals()
        print(it, "types", g.collisionType, "gpu", g.totalAmountCollisions, "orc", to["totalAmountCollisions"], "sel", g.numberSelected, "Pnull", g.P_null,
              "coll", g.totalNumberCollidableParticles, to["totalNumberCollidableParticles"])
        print("   loss/coll gpu", g.totalEnergyLoss / np.maximum(g.totalAmountCollisions, 1) / op.e / 2, "orc", to["totalEnergyLoss"] / np.maximum(to["totalAmountCollisions"], 1) / op.e / 2)
        print("   inc/coll gpu", g.totalIncidentEnergy / np.maximum(g.totalAmountCollisions, 1), "orc", to["totalIncidentEnergy"] / np.maximum(to["totalAmountCollisions"], 1))
