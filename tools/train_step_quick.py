import sys, json, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tools')
import bench_train_step as TS
out={}
for B in (128,512):
    fwd, fwd_bwd, il, cl = TS.make_step(B, "bf16", fused=True)
    with torch.no_grad():
        t_f = TS.timeit(fwd, iters=50, warm=10)
    t_fb = TS.timeit(fwd_bwd, iters=50, warm=10)
    out[B]={"fwd_ms":round(t_f,4),"fwd_bwd_ms":round(t_fb,4)}
print(json.dumps(out))
