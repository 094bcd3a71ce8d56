import os, sys, time, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dpt_oracle as O
from muggled_dpt_b200 import make_dpt_from_state_dict
t=time.time()
sd = O.giantify(O.make_synthetic_state_dict("vitg", seed=5), seed=5)
print("sd", time.time()-t, sum(v.numel() for v in sd.values())/1e9, "B params")
img = O.make_input(1, 224, 308, seed=3)
t=time.time(); ref = O.forward(sd, img, return_stages=True); print("oracle", time.time()-t)
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "depth_anything_v2_vitg.pth"); torch.save(sd, path); del sd
    cfg, model = make_dpt_from_state_dict(path)
print(cfg)
for dtype in (torch.bfloat16, torch.float16):
    model.to(device="cuda", dtype=dtype)
    with torch.inference_mode():
        x = img.to("cuda", dtype)
        d = model(x)
        tokens, grid = model.patch_embed(x); taps = model.imgencoder(tokens, grid); maps = model.reassemble(*taps, grid); fused = model.fusion(*maps)
    def err(a,b):
        a,b=a.float().cpu(),b.float().cpu(); return ((a-b).norm()/b.norm()).item(), ((a-b).abs().max()/b.abs().max()).item()
    print(dtype, "depth", err(d, ref["depth"]), "fused", err(fused, ref["fused"]), [err(taps[i], ref["taps"][i])[0] for i in range(4)], [err(maps[i], ref["maps"][i])[0] for i in range(4)])
    # timing B=8 504
    xb = torch.randn(8,3,504,504, device="cuda", dtype=dtype)
    with torch.inference_mode():
        for _ in range(3): model(xb)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): model(xb)
        e1.record(); torch.cuda.synchronize()
    print("B=8 504^2 ms/step", e0.elapsed_time(e1)/5)
