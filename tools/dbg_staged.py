"""Debug driver: one CPN tail call through the staged kernel, compared with the oracle."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases, oracle
import master_thesis_b200 as mtb
spec = dict(seed=71, b=3, f=4, h=200, w=204, sigma=float(sys.argv[1]) if len(sys.argv) > 1 else 0.1)
x, m, m_t, theta = cases.cpn_inputs(spec)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
xa, va, vm = mtb.cpn_align_tail(d(x), d(m), d(m_t), d(theta))
torch.cuda.synchronize()
oxa, ova, ovm = oracle.cpn_align_tail(x, m, m_t, theta=theta)
for n, g, o in (("xa", xa, oxa), ("va", va, ova), ("vm", vm, ovm)):
    g = g.contiguous().cpu().numpy()
    print(n, "equal" if np.array_equal(g, o) else "DIFF max %g, %d px" % (np.abs(g - o).max(), (g != o).sum()))
