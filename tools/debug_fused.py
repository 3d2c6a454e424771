import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200"), os.path.join(ROOT, "tests")]
from g2048 import update, env
from test_update_fused_gpu import _model, _boards, _samples
from helpers import ref_ppo_loss_torch
F = torch.nn.functional
h, L, n = 196, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
use_flags = len(sys.argv) > 2 and sys.argv[2] == "flags"
m = _model(h, L, 17)
boards = _boards(n, 19)
old, actions, legal, adv, g_norm = _samples(n, 23)
flags = torch.full((n,), 0x80, dtype=torch.uint8, device="cuda")
if use_flags: flags[::5] = 0
valid = flags != 0
nv = int(valid.sum())
keep = {}
m.zero_grad()
update.loss_and_grads(m, boards, actions, legal, old, adv, g_norm, flags=flags, n_total=nv, clip_eps=0.2, critic_strength=0.2, entropy_strength=0.02, keep=keep)
P = {k: v.detach().double().requires_grad_(True) for k, v in m.named_parameters()}
x48 = env.encode(boards).double()
zs, hs = [], []
z = x48 @ P["stem.0.weight"].T; z.retain_grad(); zs.append(z)
x = F.relu(F.layer_norm(z, (h,), P["stem.1.weight"], P["stem.1.bias"], 1e-5)); hs.append(x)
for l in range(L):
    pre = f"backbone.{l}.mlp."
    z = x @ P[pre + "0.weight"].T; z.retain_grad(); zs.append(z)
    x = x + F.relu(F.layer_norm(z, (h,), P[pre + "1.weight"], P[pre + "1.bias"], 1e-5)); hs.append(x)
logits = x @ P["action_head.weight"].T + P["action_head.bias"]
value = x @ P["value_head.weight"].T + P["value_head.bias"]
loss, _ = ref_ppo_loss_torch(logits[valid], value[valid], old[valid].double(), actions[valid], legal[valid], adv[valid].double(), g_norm[valid].double(), 0.2, 0.2, 0.02)
loss.backward()
def rel(a, b): return float((a.double() - b).norm() / b.norm())
for l in range(L + 1):
    print(f"h[{l}] fro {rel(keep['h_out'][l], hs[l].detach()):.2e}   dz[{l}] fro {rel(keep['dz_out'][l], zs[l].grad):.2e}")
    d = (keep['dz_out'][l].double() - zs[l].grad).abs().amax(dim=1)
    bad = torch.nonzero(d > 1e-3 * zs[l].grad.abs().max()).flatten()
    print("   bad rows:", bad[:20].tolist(), "count", bad.numel(), " invalid among bad:", int((~valid[bad]).sum()))
for k, p in m.named_parameters():
    print(f"{k}: fro {rel(p.grad, P[k].grad):.2e}")
# ---- the worst row in detail, and the same graph in float32 torch
r = int((keep['dz_out'][L].double() - zs[L].grad).abs().amax(dim=1).argmax())
print("row", r, "valid", bool(valid[r]), "adv", float(adv[r]), "legal", int(legal[r]), "action", int(actions[r]))
print("ours  dz[L][r][:6]", keep['dz_out'][L][r][:6].tolist())
print("ref64 dz[L][r][:6]", zs[L].grad[r][:6].tolist())
print("dhead ours", keep['dhead'][r].tolist())
lg = logits.detach()[r]; print("logits64", lg.tolist(), "old", old[r].tolist())
ratio_in = torch.log_softmax(torch.where(((legal[r].long() >> torch.arange(4, device='cuda')) & 1) == 1, lg, torch.tensor(float('-inf'), device='cuda', dtype=torch.float64)), -1)[actions[r].long()] - old[r][actions[r].long()]
print("log ratio", float(ratio_in), "ratio", float(ratio_in.exp()))
zL = zs[L].detach()[r]; print("z row mean/std", float(zL.mean()), float(zL.std()))
