import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import latentdiffeq_jl_b200 as ldeq
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "c2_latentode_mlp.npz"))
dev = "cuda:0"
dims = g["dims"].tolist()
for env in ("", "1"):
    if env: os.environ["LDEQ_MLP_NO_RESIDENT"] = "1"
    tr, st, _ = ldeq.mlp_solve_raw(torch.from_numpy(g["z0"]).to(dev), torch.from_numpy(g["params"]).to(dev), dims, g["t"])
    ref = g["traj_f32_adaptive_global"]
    print("resident" if not env else "general", "naccept", st.naccept.unique().tolist(), "nreject", st.nreject.unique().tolist(), "golden", int(g["naccept_f32"]),
          "max rel err", float(np.abs(tr.cpu().numpy() - ref).max() / np.abs(ref).max()))
