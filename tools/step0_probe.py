#!/usr/bin/env python
"""Step 0 of an image (FeatureNeRF in all 12 pose blocks + one guided step), replayed from its CUDA graph, with
and without the row-class sharing of the FeatureNeRF encoding (FusedGuidedStep.dedup_rows), A/B in one process:
1 image (UNet batch 3: 2 classes) and a sweep unit of 4 prompts on one target camera (UNet batch 12: 2 classes).

    python tools/step0_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from custom_diffusion360_b200 import synthetic as S
from custom_diffusion360_b200.sgm.modules.diffusionmodules.sampling import FusedGuidedStep

dev = torch.device("cuda:0")
engine, net, step, x_init, sigmas = bench.make_step(128, 1, dev, 0, True)
del step
cfg = dict(S.SDXL_CFG)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for n_img in (1, 4):
        cond, uc = S.make_conditioning(cfg, n_img, dev, seed=17)
        poses = [S.lookat_cameras(8, seed=3, target_azimuth=0.35)] * n_img
        x0 = torch.randn(n_img, 4, 128, 128, device=dev) * float(torch.sqrt(1.0 + sigmas[0] ** 2))
        for dedup in (False, True):
            net.clear_rendered_feat()
            st = FusedGuidedStep(net, engine.denoiser, engine.sampler.guider, cond, uc, pose=poses, n_img=n_img,
                                 latent_shape=(4, 128, 128), dedup_rows=dedup)
            x = x0.clone()
            times = []
            for image in range(4):      # image 0 eager, image 1 captures the step-0 graph, images 2, 3 replay it
                net.clear_rendered_feat()
                torch.cuda.synchronize()
                e0.record()
                st(x, float(sigmas[0]), float(sigmas[1]))
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
                for i in range(1, 4):
                    st(x, float(sigmas[i]), float(sigmas[i + 1]))
            torch.cuda.synchronize()
            e0.record()
            for i in range(10):
                st(x, float(sigmas[5 + i]), float(sigmas[6 + i]))
            e1.record()
            torch.cuda.synchronize()
            print(f"n_img={n_img} dedup_rows={dedup}: classes={None if st._class_key is None else len(st._class_key[0])} "
                  f"step0 replay {times[2]:.1f} / {times[3]:.1f} ms (eager {times[0]:.1f}), steady step {e0.elapsed_time(e1) / 10:.2f} ms")
            del st
