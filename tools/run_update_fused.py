"""One fused update call (forward + loss + backward-data + weight gradients) on N random samples: the ncu target."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2048-ppo_b200"), os.path.join(ROOT, "tests")]
from g2048 import update
from test_update_fused_gpu import _model, _boards, _samples

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
drop = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0      # Dropout(p) of the update forward (the reference's model: 0.1)
m = _model(196, 2, 1)
boards = _boards(n, 2)
old, actions, legal, adv, g_norm = _samples(n, 3)
packed = update.pack(m)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    m.zero_grad()
    e0.record()
    update.loss_and_grads(m, boards, actions, legal, old, adv, g_norm, packed=packed, dropout_p=drop, dropout_seed=7)
    e1.record()
    torch.cuda.synchronize()
    print(f"n={n} dropout={drop}: {e0.elapsed_time(e1):.3f} ms")
