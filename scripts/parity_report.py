"""Dev tool: a fixture (argv[3], default room), B200 engine vs live reference CPU backend, per-step drift."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle.runner import run_reference
from sims import SIMS
from trace_utils import make_inputs, rollout_gpu

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 250
sim = sys.argv[3] if len(sys.argv) > 3 else "room"
cfg = {"episode_len": 100, "seed": 1}
ins = make_inputs(sim, W, steps, seed=5)
ref, _ = run_reference(SIMS[sim], W, steps, ins, cfg, workers=1)
got, nk = rollout_gpu(sim, W, steps, ins, cfg)
print("kernels per step", nk)
for k in ref:
    r, g = ref[k], got[k]
    if isinstance(r, list):
        bad_shape = [t for t in range(len(r)) if r[t].shape != g[t].shape]
        if bad_shape:
            print(k, "shape mismatch at steps", bad_shape[:5]); continue
        if np.issubdtype(r[0].dtype, np.integer):
            bad = [t for t in range(len(r)) if not np.array_equal(r[t], g[t])]
            print(k, "int mismatches at steps:", bad[:10])
        else:
            errs = [float(np.abs(r[t] - g[t]).max()) if len(r[t]) else 0.0 for t in range(len(r))]
            first = next((t for t, e in enumerate(errs) if e > 0), None)
            print(k, "max abs err", max(errs), "first nonzero step", first,
                  "err@[1,10,50,99,150,249]", [errs[min(t, len(errs)-1)] for t in (1, 10, 50, 99, 150, 249)])
    else:
        if np.issubdtype(r.dtype, np.integer):
            bad = np.argwhere(r != g)
            print(k, "int mismatches:", len(bad), bad[:3].tolist())
        else:
            err = np.abs(r - g).reshape(r.shape[0], -1).max(axis=1)
            first = next((t for t, e in enumerate(err) if e > 0), None)
            print(k, "max abs err", float(err.max()), "first nonzero step", first)
            if first is not None:
                idx = np.argwhere(r[first] != g[first])[:4]
                for i in idx:
                    print("    step", first, "at", i.tolist(), "ref", r[first][tuple(i)], "got", g[first][tuple(i)])
