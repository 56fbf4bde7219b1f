"""tests/fuzz/fuzz_raster.py -- TEST INFRASTRUCTURE, run by hand (not collected by pytest), emulated build only.

Both rasterisation rules on lattice polygons (tests/checks.py: adversarial_polygon) against the oracle.

    python tests/fuzz/fuzz_raster.py <first seed> <number of seeds>     (prints one FAIL line per seed that differs)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from imageprocess_b200.ops import Engine
from imageprocess_b200 import geometry as geo
from tests.emu.emu_backend import NumpyMem, emu_lib
from tests import checks
from oracle import port, shims
import time
eng = Engine(emu_lib(), NumpyMem())
H,W=24,40
t0=time.time()
nbad=0
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for seed in range(seed0, seed0 + n_seeds):
    rng=np.random.default_rng(seed+5000)
    polys=[checks.adversarial_polygon(rng,H,W) for _ in range(300)]
    specs=[geo.mpl_spec(P,(W,H)) for P in polys]
    rm=eng.rasterize(geo.RULE_MPL,specs,(H,W),1,want_union=False)
    for i,P in enumerate(polys):
        want=port.rasterize_polygon(P,(H,W)); x0,y0,x1,y1=specs[i].srect
        got=np.zeros((H,W),bool); got[y0:y1,x0:x1]=rm.mask_host(i)
        if (got^want).sum(): nbad+=1; print("mpl",seed,i,P.tolist())
    specs=[geo.sk_spec(P[:,1],P[:,0],(H,W)) for P in polys]
    rm=eng.rasterize(geo.RULE_SK,specs,(H,W),1,want_union=False)
    for i,P in enumerate(polys):
        want=np.zeros((H,W),bool); rr,cc=shims.polygon(P[:,1],P[:,0],(H,W)); want[rr,cc]=True
        if (rm.mask_host(i)^want).sum(): nbad+=1; print("sk",seed,i,P.tolist())
print("bad", nbad, "of", n_seeds * 300 * 2, round(time.time() - t0, 1))
