#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x -k "narrower or golden or tiny" > gpurun_out/n_pytest.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/n_pytest.log
python - <<'PY'
import torch, time
from nerf_sr_b200 import Renderer, synthetic as S
for W, prec in ((128, "bf16x3"), (128, "fp32_simt"), (256, "bf16x3")):
    cfg = S.RenderConfig(W=W, white_bkgd=True)
    r = Renderer(cfg, torch.device("cuda:0"), precision=prec)
    r.load_state_dict(0, S.make_mlp_params(cfg, 4)); r.load_state_dict(1, S.make_mlp_params(cfg, 17))
    rays = S.synthetic_rays(160000, 7, "blender").cuda()
    for _ in range(2): r.forward_rays(rays, want_weights=False)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r.forward_rays(rays, want_weights=False); r.forward_rays(rays, want_weights=False); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    print(f"W={W} {prec}: {ms:.1f} ms per 160000-ray frame = {160000/ms/1e3:.3f} M rays/s", flush=True)
    r.close()
PY
