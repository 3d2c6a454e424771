import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200"), os.path.join(ROOT, "tests")]
from g2048 import update, env, fused, ppo
from test_train_cpu import policy_b
from helpers import ref_ppo_loss_torch
F = torch.nn.functional
G = lambda n: np.load(os.path.join(ROOT, "tests", "golden", n + ".npz"))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
gr, ga, gl = G("rollout"), G("advantage"), G("loss")
m = policy_b(G).cuda(); m.train()
boards = cu(gr["board"].view(np.int64)); n = boards.numel()
adv = cu(ga["readme__adv"].astype(np.float32)); gn = cu(ga["readme__g_norm"].astype(np.float32))
ent, crit = gl["readme__coef"].tolist()
old, actions, legal = cu(gr["logp"]), cu(gr["action"]), cu(gr["legal"])
h, L = 192, 2
keep = {}
m.zero_grad()
update.loss_and_grads(m, boards, actions, legal, old, adv, gn, clip_eps=0.2, critic_strength=crit, entropy_strength=ent, keep=keep)
ours = {k: p.grad.clone() for k, p in m.named_parameters()}
P = {k: v.detach().double().requires_grad_(True) for k, v in m.named_parameters()}
x48 = env.encode(boards).double()
zs, ys = [], []
z = x48 @ P["stem.0.weight"].T; z.retain_grad(); zs.append(z)
y = F.layer_norm(z, (h,), P["stem.1.weight"], P["stem.1.bias"], 1e-5); ys.append(y); x = F.relu(y)
for l in range(L):
    pre = f"backbone.{l}.mlp."
    z = x @ P[pre + "0.weight"].T; z.retain_grad(); zs.append(z)
    y = F.layer_norm(z, (h,), P[pre + "1.weight"], P[pre + "1.bias"], 1e-5); ys.append(y); x = x + F.relu(y)
logits = x @ P["action_head.weight"].T + P["action_head.bias"]
value = x @ P["value_head.weight"].T + P["value_head.bias"]
loss, _ = ref_ppo_loss_torch(logits, value, old.double(), actions, legal, adv.double(), gn.double(), 0.2, crit, ent)
loss.backward()
rel = lambda a, b: float((a.double() - b).norm() / b.norm())
for l in range(L + 1):
    d = (keep['dz_out'][l].double() - zs[l].grad).abs().amax(dim=1)
    bad = torch.nonzero(d > 1e-3 * zs[l].grad.abs().max()).flatten()
    print(f"dz[{l}] fro {rel(keep['dz_out'][l], zs[l].grad):.2e} bad rows {bad.tolist()[:12]} ; min|y| in those rows {[float(ys[l][b].abs().min()) for b in bad[:6]]}")
    print("   smallest |y| overall:", torch.sort(ys[l].abs().flatten())[0][:5].tolist())
for k in ours:
    ref_np = gl["readme__grad__" + k.replace(".", "__")]
    # the fixture holds CLIPPED grads: compare directions
    a, b = ours[k].double().cpu().flatten(), torch.from_numpy(ref_np).double().flatten()
    print(f"{k}: vs fp64 autograd fro {rel(ours[k], P[k].grad):.2e}; cos vs fixture {float((a @ b) / (a.norm() * b.norm())):.8f}")
# the x3 autograd path and the cuBLAS fp32 path on the same data
for mm in ("x3", "cublas"):
    m.zero_grad()
    lo, v = fused.mlp_forward(m, env.encode(boards), matmul=mm)
    ls, st = ppo.ppo_loss(lo, v, old, actions, legal, adv, gn, clip_eps=0.2, critic_strength=crit, entropy_strength=ent)
    ls.backward()
    print(mm, "autograd path vs fp64:", {k: f"{rel(p.grad, P[k].grad):.1e}" for k, p in m.named_parameters() if "0.weight" in k})
