"""Where the fp32 error of the fusion module comes from: per-stage max |x - fp64| of the drop-in and of the reference module
(fp32, same GPU), the fp64 run of the reference module being the yardstick (diagnostic tool)."""
import copy, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import diffreg_b200
from oracle import ref_loader

torch.backends.cuda.matmul.allow_tf32 = False
ref = ref_loader.load_fusion()
g = torch.Generator().manual_seed(62)
blocks = ["self", "cross"] * 3
rnet = ref.fusion.CrossModalFusionModule(512, 512, 256, 256, 4, blocks).eval()
if len(sys.argv) > 1 and sys.argv[1] == "random":
    with torch.no_grad():
        for name, prm in rnet.named_parameters():
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.3 if prm.dim() == 1 else 1.0 / prm.shape[-1] ** 0.5))
            if name.endswith("norm.weight"):
                prm.add_(1.0)
rnet = rnet.cuda()
r64 = copy.deepcopy(rnet).double()
onet = diffreg_b200.CrossModalFusionModule(512, 512, 256, 256, 4, blocks).cuda().eval()
onet.load_state_dict(rnet.state_dict(), strict=True)
n_img, n_pcd = 2048, 4800
img, dino = torch.randn(1, n_img, 512, generator=g).cuda(), torch.randn(1, n_img, 1024, generator=g).cuda()
pcd = torch.randn(1, n_pcd, 512, generator=g).cuda()
pix = (torch.rand(1, n_img, 2, generator=g) * 2.0 - 1.0).cuda()
pts = (torch.randn(1, n_pcd, 3, generator=g) * 0.8 + torch.tensor([0.3, -0.2, 2.0])).cuda()


def trace(net, dbl):
    rec = []
    hooks = [net.transformer[i].register_forward_hook(lambda m, a, o, i=i: rec.append((f"block{i}", o.detach().double()))) for i in range(6)]
    c = (lambda t: t.double()) if dbl else (lambda t: t)
    with torch.no_grad():
        e2 = net.create_2d_embedding(c(pix)); e3 = net.create_3d_embedding(c(pts))
        out = net(c(img), c(dino), c(pix), c(pcd), c(pts))
    for h in hooks:
        h.remove()
    return [("emb2d", e2.double()), ("emb3d", e3.double())] + rec + [("img_out", out[0].double()), ("pcd_out", out[1].double())]


t64, tr, to = trace(r64, True), trace(rnet, False), trace(onet, False)
for (n, a), (_, b), (_, c) in zip(t64, tr, to):
    print(json.dumps({"stage": n, "ref32_err": (b - a).abs().max().item(), "dropin_err": (c - a).abs().max().item(),
                      "dropin_vs_ref32": (c - b).abs().max().item(), "scale": a.abs().max().item()}))
