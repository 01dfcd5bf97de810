"""Development: bias-map path vs per-read pool-add (DAN_B200_NO_BMAP is read once per process -> run twice), batch-split consistency."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dl4vc_b200.config import small_config
from dl4vc_b200.factory import build_model
from dl4vc_b200.synth import make_pileups
from dl4vc_b200.weights import synth_state_dict

cfg = small_config(**({'highway': False} if os.environ.get('NOHW') else {}))
sd = synth_state_dict(cfg, seed=9)
base = make_pileups(23, seed=11, coverage="poisson")
def heads(model, arrs):
    t = [torch.from_numpy(np.ascontiguousarray(a)) for a in arrs]
    r, q, s, ref, rm, vm = t
    return model.forward_heads(r, ref, q, s, rm, vm).cpu().numpy()
m = build_model(cfg, sd, precision="fp32")
ref32 = heads(m, base.arrays())
m.set_precision("bf16")
full = heads(m, base.arrays())
parts = np.concatenate([heads(m, base.slice(lo, min(lo + 10, 23)).arrays()) for lo in range(0, 23, 10)])
scale = np.abs(ref32).max()
print("mode", "NO_BMAP" if os.environ.get("DAN_B200_NO_BMAP") else "BMAP")
print(" bf16(23) vs fp32:", np.abs(full - ref32).max() / scale)
print(" bf16(10+10+3) vs fp32:", np.abs(parts - ref32).max() / scale)
print(" bf16(23) vs bf16(10+10+3):", np.abs(full - parts).max() / scale, "per-candidate max:", (np.abs(full - parts).max(1) / scale).round(5))
np.save("gpurun_out/diag_%s.npy" % ("nobmap" if os.environ.get("DAN_B200_NO_BMAP") else "bmap"), full)
again = heads(m, base.arrays())
print(" run-to-run (same 23):", np.abs(full - again).max() / scale, "candidates differing:", np.nonzero(np.abs(full - again).max(1))[0])
p2 = np.concatenate([heads(m, base.slice(lo, min(lo + 10, 23)).arrays()) for lo in range(0, 23, 10)])
print(" run-to-run (10+10+3):", np.abs(parts - p2).max() / scale, "candidates differing:", np.nonzero(np.abs(parts - p2).max(1))[0])
big = make_pileups(296, seed=5, coverage="poisson")
a = heads(m, big.arrays()); b = heads(m, big.arrays())
print(" run-to-run (296):", np.abs(a - b).max() / np.abs(a).max(), "n differing:", len(np.nonzero(np.abs(a - b).max(1))[0]))
