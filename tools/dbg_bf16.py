"""Debug aid: prediction / gradient distance between precision modes (fp32, bf16 activations + CUDA-core convs, bf16 tcgen05)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import awr_oracle as O
import awr_b200

net = sys.argv[1] if len(sys.argv) > 1 else "resnet_18"
B, H, J, ds = int(sys.argv[2]) if len(sys.argv) > 2 else 2, 128, 14, 2
kind, n = net.split("_")
ks = 1.0 if kind == "resnet" else 0.4
sd = O.randomize_bn(O.resnet_deconv_init(int(n), J, ds, 21, head_std=0.02), 22) if kind == "resnet" else O.randomize_bn(O.hourglass_init(int(n), J, 21), 22)
img, jt = O.synthetic_batch(B, H, J, 23)
img, jt = img.cuda(), jt.cuda()
FM, crit = awr_b200.FeatureModule(), awr_b200.My_SmoothL1Loss()
res = {}
for name, prec, simt in [("fp32", "fp32", "0"), ("bf16_simt", "bf16", "1"), ("bf16_tc", "bf16", "0")]:
    os.environ["AWR_B200_DEBUG_SIMT"] = simt
    m = awr_b200.get_deconv_net(int(n), J, ds, precision=prec) if kind == "resnet" else awr_b200.PoseNet(net, J, precision=prec)
    m.load_state_dict(sd); m = m.cuda().train()
    gt = FM.joint2offset(jt, img, ks, H // ds)
    o = m(img); pred = o[-1] if isinstance(o, list) else o
    uvd = FM.offset2joint_softmax(pred, img, ks)
    lc, ld = crit(uvd, jt), crit(pred, gt)
    m.zero_grad(); (lc + ld).backward()
    res[name] = (pred.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, lc.item(), ld.item())
for a, b in [("fp32", "bf16_simt"), ("fp32", "bf16_tc"), ("bf16_simt", "bf16_tc")]:
    pa, ga, *la = res[a]; pb, gb, *lb = res[b]
    print(f"{a} vs {b}: pred relL2 {(pa - pb).norm().item() / pa.norm().item():.4f} max {(pa - pb).abs().max().item():.4f}  loss {la} {lb}")
    worst = []
    for k in ga:
        if ga[k].numel() < 64 or ga[k].abs().mean() < 1e-7: continue
        cos = torch.nn.functional.cosine_similarity(ga[k].flatten().double(), gb[k].flatten().double(), dim=0).item()
        worst.append((cos, gb[k].norm().item() / ga[k].norm().item(), k))
    keys = [k for k in ga if ga[k].numel() >= 64 and ga[k].abs().mean() >= 1e-7]
    fa = torch.cat([ga[k].flatten().double() for k in keys]); fb = torch.cat([gb[k].flatten().double() for k in keys])
    print("    flat grad cosine %.4f  norm ratio %.4f" % (torch.nn.functional.cosine_similarity(fa, fb, dim=0).item(), fb.norm().item() / fa.norm().item()))
    big = sorted(keys, key=lambda k: -ga[k].norm().item())[:5]
    for k in big:
        print("    big: %-40s norm %.3e cos %.4f" % (k, ga[k].norm().item(), torch.nn.functional.cosine_similarity(ga[k].flatten().double(), gb[k].flatten().double(), dim=0).item()))
    worst.sort()
    for w in worst[:6]: print("    cos %.4f ratio %.4f %s" % w)

# reference point: the oracle's functional torch model on the GPU, fp32 vs torch.autocast(bf16)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sdc = {k: v.cuda() for k, v in sd.items()}
outs = {}
for name, ac in [("torch_fp32", False), ("torch_autocast_bf16", True)]:
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
        loss, lc, ld, uvd, pred, grads, _ = O.loss_and_grads(sdc, img, jt, net, ds, ks, 1.0, 1.0)
    outs[name] = (pred.float(), {k: g.float() for k, g in grads.items() if g is not None})
for a, b in [("torch_fp32", "torch_autocast_bf16")]:
    pa, ga = outs[a]; pb, gb = outs[b]
    keys = [k for k in ga if ga[k].numel() >= 64 and ga[k].abs().mean() >= 1e-7]
    fa = torch.cat([ga[k].flatten().double() for k in keys]); fb = torch.cat([gb[k].flatten().double() for k in keys])
    print(f"{a} vs {b}: pred relL2 {(pa - pb).norm().item() / pa.norm().item():.4f}  flat grad cosine "
          f"{torch.nn.functional.cosine_similarity(fa, fb, dim=0).item():.4f}")
pa, ga = outs["torch_fp32"]; pb, gb, *_ = res["fp32"]
keys = [k for k in ga if ga[k].numel() >= 64 and ga[k].abs().mean() >= 1e-7]
fa = torch.cat([ga[k].flatten().double() for k in keys]); fb = torch.cat([gb[k].flatten().double() for k in keys])
print(f"torch_fp32 vs ours fp32: pred relL2 {(pa - pb).norm().item() / pa.norm().item():.2e} flat grad cosine {torch.nn.functional.cosine_similarity(fa, fb, dim=0).item():.6f}")
