"""Full-width UNet forward at the configs[0] shape run several times on identical inputs: per-tap bitwise comparison with the
first run, then the public forward.  Prints the first tap that differs (a race or an uninitialised read would show here)."""
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from test_unet_gpu import build_product
from oracle import vx_oracle as O
cfg = O.DEFAULT_CFG
sd = O.synth_state_dict(O.unet_param_shapes(cfg), 1234)
lat, kps, audio, banks = O.synth_inputs(cfg, 4, 64, 64, True, 42)
model, reader = build_product(cfg, sd, [b[1:] for b in banks], 0.95, 3.0)
x = lat.repeat(2, 1, 1, 1, 1)
enc = audio.reshape(-1, 5, cfg["cross_attention_dim"])
eng = model.engine()
b, c, f, h, w = x.shape
frames = x.bfloat16().cuda().permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w).contiguous()
kps_nhwc = kps.bfloat16().cuda().permute(0, 2, 3, 4, 1).reshape(b * f * h * w, -1).contiguous()
first = None
for run in range(4):
    taps = {}
    out = eng.forward_frames(frames, 499, enc.cuda(), kps_nhwc, None, b, f, taps=taps)
    torch.cuda.synchronize()
    taps["out"] = out.float()
    if first is None:
        first = {k: v.clone() for k, v in taps.items()}
        continue
    bad = [(k, (taps[k] - first[k]).abs().max().item(), (taps[k] != first[k]).float().mean().item()) for k in first if not torch.equal(taps[k], first[k])]
    print(f"run {run}: {len(bad)} / {len(first)} taps differ from run 0; first: {bad[:3]}", flush=True)
out_nt = eng.forward_frames(frames, 499, enc.cuda(), kps_nhwc, None, b, f)
print("no-taps forward_frames == run 0:", torch.equal(out_nt.float(), first["out"]), (out_nt.float() - first["out"]).abs().max().item())
out2 = model(x.bfloat16().cuda(), 499, encoder_hidden_states=enc.bfloat16().cuda(), kps_features=kps.bfloat16().cuda(), return_dict=False)[0]
o0 = first["out"].view(b, f, -1, h, w).permute(0, 2, 1, 3, 4)
print("public forward == run 0:", torch.equal(out2.float(), o0), (out2.float() - o0).abs().max().item(), (out2.float() != o0).float().mean().item())
