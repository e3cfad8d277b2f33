import sys, time, torch
sys.path.insert(0, '.')
from types import SimpleNamespace
from deltakd_b200 import functional as Fn, synth, heads as H
from oracle import losses as O

def run(B, ratio, mode="bf16x3"):
    Fn.set_matmul_precision(mode)
    args = SimpleNamespace(distillation_type="mgd")
    teacher, student = synth.FeatureReplayModel(384), synth.FeatureReplayModel(192)
    torch.manual_seed(1)
    H.attach_distillation_heads(student, teacher, args)
    with torch.no_grad():
        student.mask_token.normal_(0, 0.1)
    s_feats, t_feats = synth.make_features(B, 3, layers=[11])
    noise = synth.make_noise(B, seed=B)
    heads64 = {k: v.detach().double().requires_grad_(True) for k, v in H.head_tensors(student).items()}
    s64 = s_feats[11].double().requires_grad_(True)
    t0 = time.time()
    ref = O.mgd([s64], [t_feats[11].double()], heads64, 7e-5, ratio, noise)
    ref.backward()
    t1 = time.time()
    student = student.cuda()
    sc = s_feats[11].cuda().requires_grad_(True)
    loss = Fn.masked_generation_loss(sc, t_feats[11].cuda(), student.align, student.mask_token, student.generation,
                                     mask_ratio=ratio, noise=noise.cuda(), scale=7e-5 / (B * 196 * 384))
    loss.backward()
    torch.cuda.synchronize()
    print(f"B={B} ratio={ratio} {mode}: loss {loss.item():.9e} ref {ref.item():.9e} oracle_time {t1-t0:.1f}s")
    g, r = sc.grad.cpu().double(), s64.grad
    print("  g_s rel", float((g - r).norm() / r.norm()))
    per_row = (g - r).norm(dim=-1) / (r.norm(dim=-1) + 1e-30)
    for b in range(B):
        bad = (per_row[b] > 1e-3).nonzero().flatten().tolist()
        print(f"  sample {b}: rows with err>1e-3: {len(bad)} first {bad[:20]} max {float(per_row[b].max()):.3e}")
    for k, p in H.head_tensors(student).items():
        print("  ", k, float((p.grad.cpu().double() - heads64[k].grad).norm() / heads64[k].grad.norm()))

for B, ratio in [(1, 0.5), (2, 0.5), (5, 0.25)]:
    run(B, ratio)
