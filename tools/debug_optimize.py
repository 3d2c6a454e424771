"""How much of the weight-update difference vs the reference run is the x3 GEMM precision?  Same loop with
torch autograd + cuBLAS fp32 (engine=cublas) or the x3 GEMM kernels under autograd (engine=x3)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200"), os.path.join(ROOT, "tests")]
from g2048 import optimize, update, fused, ppo, env
from test_train_cpu import policy_b
from test_optimize_gpu import _episodes, SgdStack
G = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
fx = G("optimize")
for engine in ("fused", "x3", "cublas"):
    m = policy_b(G).cuda()
    init = {k: v.detach().clone() for k, v in m.state_dict().items()}
    batch = optimize.episodes_to_batch(_episodes(G), torch.device("cuda"))
    opt = SgdStack(m, float(fx["lr"]))
    torch.manual_seed(int(fx["seed"]))
    n = batch["boards"].numel()
    for _ in range(int(fx["epochs"])):
        order = optimize._epoch_order(n).cuda()
        for i in range(0, n, int(fx["batch_size"])):
            b = {k: v[order[i:i + int(fx["batch_size"])]] for k, v in batch.items()}
            m.train()
            if engine == "fused":
                update.loss_and_grads(m, b["boards"], b["actions"], b["legal"], b["logp"], b["adv"], b["g_norm"], clip_eps=0.2,
                                      critic_strength=0.2, entropy_strength=0.02)
            else:
                lo, v = fused.mlp_forward(m, env.encode(b["boards"]), matmul=engine)
                loss, _ = ppo.ppo_loss(lo, v, b["logp"], b["actions"], b["legal"], b["adv"], b["g_norm"], clip_eps=0.2,
                                       critic_strength=0.2, entropy_strength=0.02)
                loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
            opt.step(); opt.zero_grad()
    out = {}
    for k, v in m.state_dict().items():
        ref = torch.from_numpy(fx["final__" + k.replace(".", "__")]).cuda()
        out[k] = float(((v - init[k]) - (ref - init[k])).norm() / (ref - init[k]).norm())
    print(engine, {k: f"{v:.1e}" for k, v in out.items()})
