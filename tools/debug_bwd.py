import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from robotoc_b200 import ANYMAL, Layout, RiccatiRecursion
from robotoc_b200.schedule import plain_schedule
from robotoc_b200.synth import make_kkt
dims = ANYMAL; L = Layout(dims)
ctrl = plain_schedule(4, 0.03, 12)
kkt, dx0 = make_kkt(dims, L, ctrl, 2, 1)
rr = RiccatiRecursion(dims, len(ctrl), 2); rr.setTimeDiscretization(ctrl)
print("created", flush=True)
rr.backwardRiccatiRecursion(kkt); rr.synchronize(); print("bwd returned", flush=True)
