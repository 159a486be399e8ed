import sys, numpy as np
sys.path.insert(0, '.')
import stochquant_b200 as sq
which = sys.argv[1]
if which == 'c3':
    g = sq.Context((64, 64, 64, 64), real="f32", math=sys.argv[2], seed=1242608872)
    g.step(0.01, int(sys.argv[3]))
    print('ok', g.measure()["seed"], g.measure()["nevents"])
else:
    g = sq.Context((256, 256, 256, 16), real="f32", math="fast")
    g.step(0.01, int(sys.argv[3]))
    print('ok', g.measure()["seed"], g.measure()["nevents"])
