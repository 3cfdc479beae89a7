"""BASELINE config 5 (evaluate.py-style multimodal inference sweep): 1 input x 64 sampled z through G_A_B
(model.generate_multi, model.py:687-696) plus the E_B forward (predict_enc_params, model.py:653-662) on [64,6,64,64];
forward-only, CUDA-event timed, against the same ops of the oracle networks (cuDNN) on the same GPU.
usage: python tools/infer_sweep.py [--reps 50]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dtg  # noqa: E402,F401
from dtg_b200 import engine, model as dmodel  # noqa: E402
from oracle import nets as onets, step as ostep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
engine.set_precision(args.precision)
opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
state = onets.init_model_state(seed=1234)
m = dmodel.AugmentedCycleGAN(opt, testing=True)
for name, net in m._nets().items():
    net.load_state_dict({k: v.clone() for k, v in state[name].items()}, strict=False)
m.prepare()
a, b, _ = [t.cuda() for t in ostep.synthetic_batch(64, seed=1)]
real_A = a[:1].contiguous()
zs = torch.randn(64, 16, 1, 1, device="cuda")


def ours():
    with torch.no_grad():
        out = m.generate_multi(real_A, zs)
        mu = m.predict_enc_params(a, b)
    return out, mu


om = ostep.OracleModel(ostep.default_opt(), state, device="cuda")


def ref():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = om.G_A_B(real_A.repeat(64, 1, 1, 1), zs)
        mu = om.E_B(torch.cat((a, b), 1))
    return out, mu


def timed(f):
    for _ in range(5):
        f()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.graph(g, stream=side):
            f()
        run = g.replay
        mode = "cuda graph"
    except Exception:
        run, mode = f, "eager"
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps, mode


torch.backends.cudnn.benchmark = True
ms_o, mode_o = timed(ours)
ms_r, mode_r = timed(ref)
o, r = ours()[0], ref()[0].float()
print(json.dumps({"config": "1 input x 64 z through G_A_B + E_B forward on [64,6,64,64], %s" % args.precision,
                  "ours_ms": ms_o, "ours_images_per_s": 64 / ms_o * 1e3, "ours_mode": mode_o,
                  "cudnn_bf16_autocast_ms": ms_r, "cudnn_images_per_s": 64 / ms_r * 1e3, "cudnn_mode": mode_r,
                  "rel_diff_vs_cudnn_bf16": float((o - r).norm() / r.norm())}))
