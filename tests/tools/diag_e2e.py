import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from helpers import load, sub
from test_gpu_e2e_head import make_trainer, engine_flags, COLS
from multi_modal_normative_modeling_b200 import _lib
g = load("tests/golden", sys.argv[1] if len(sys.argv) > 1 else "e2e_M3_full")
for engine in ("tcs", "fp32"):
    for steps in (1, 2, g["eps"].shape[0]):
        tr, _ = make_trainer(g, keep_grads=False)
        losses = tr.train_steps(steps, eps=torch.from_numpy(g["eps"][:steps]).cuda()[None], record_losses=True, flags=engine_flags(engine) | _lib.TRAIN_LOSS8)
        torch.cuda.synchronize()
        print(engine, steps, "loss rel", np.abs(losses[0].cpu().numpy()[:, list(COLS)] / g["losses"][:steps] - 1).max(0))
        if steps == g["eps"].shape[0]:
            sd, init = tr.state_dict(0), sub(g, "init/")
            for k, v in sub(g, "final/").items():
                if "running" in k or "num_batches" in k: continue
                d_ref, d_got = v - init[k], sd[k].cpu().numpy().reshape(v.shape) - init[k]
                dev = np.sort((np.abs(d_got - d_ref) / np.abs(d_ref).max()).ravel())
                if dev[min(dev.size - 1, int(dev.size * .99))] > 1e-3: print("  %-50s n %6d  q50 %.1e q90 %.1e q99 %.1e max %.1e" % (k, dev.size, dev[dev.size // 2], dev[int(dev.size * .9)], dev[min(dev.size - 1, int(dev.size * .99))], dev[-1]))
        tr.close()
