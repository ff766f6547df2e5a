"""Stage-by-stage comparison of the row-tiled path (world 1) against the untiled model (debugging aid)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from dualpixelface_b200.synthetic import synthetic_batch
from dualpixelface_b200.tiled import HALO, TiledStereoDPNet
from test_gpu_models import build, calibrated_state

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (192, 160)
if "bench" in sys.argv:                                         # the bench's model (seeded synthetic 'calibrated' weights)
    from bench import build_model
    model = build_model(torch.device("cuda"))
else:
    st, _ = calibrated_state("stereodpnet", synthetic_batch(2, 128, 160, training=True, seed=0))
    model = build("stereodpnet"); model.load_state_dict(st, strict=False); model.cuda().eval()
model.encoder_autocast = "bf16enc" in sys.argv
batch = {k: v.cuda() for k, v in synthetic_batch(1, h, w, training=True, seed=2).items()}
def err(a, b):
    d = (a.float() - b.float()).abs()
    return f"max {float(d.max()):.5f} mean {float(d.mean()):.6f} (ref max {float(b.float().abs().max()):.3f}, mean |ref| {float(b.float().abs().mean()):.4f})"
with torch.no_grad():
    ref_img, tgt_img = model._select_views(batch)
    fr, ft = model._features(ref_img, tgt_img)
    vol = model.cost_volume(fr, ft)
    cost_i, outs = model.aggregation(vol)
    disp, _ = model.regression_layer(cost_i)
    normals, _, _ = model.normal_estimator([outs[0]], [disp[0]], batch)
    tm = TiledStereoDPNet(model, h, 0, 1)
    x = torch.cat([ref_img, tgt_img], 0)
    f = tm.enc(x).to(torch.bfloat16)
    print("features ref", err(f[:1], fr), " tgt", err(f[1:], ft))
    # feed the UNTILED features so that later stages are compared on identical inputs
    v = tm._volume(fr.contiguous(), ft.contiguous())
    print("volume", err(v[:, :, HALO:-HALO], vol))
    c3, o3 = tm._aggregate(tm._ext(vol))
    print("cost3", err(c3[:, :, HALO:-HALO], cost_i[0]), " out3", err(o3[:, :, HALO:-HALO], outs[0]))
    d = tm._regress(tm.t.halo_cat(cost_i[0].unsqueeze(-1), HALO, HALO, 2).squeeze(-1).contiguous())
    print("disp (same cost)", err(d, disp[0]))
    n = tm._normals(tm._ext(outs[0]), disp[0], batch)
    print("normal (same inputs)", err(n, normals[0]))
    full = tm(batch)
    print("end to end: disp", err(full["pred_depth"][:, 0], disp[0]), " normal", err(full["pred_normal"][:, 0], normals[0]))
