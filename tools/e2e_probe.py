#!/usr/bin/env python
"""Where the end-to-end (host-buffer) step loses time against the device-resident one: back-to-back
replays, replay + synchronise, FusedGuidedStep.__call__ + synchronise, step_host (H2D + step + D2H +
synchronise), host cost of the scalar refresh and of graph.replay().  Found the 1.4 ms per call spent
walking the module tree (`net.pose_blocks()`) that only shows when every step synchronises."""
import sys, time, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
engine, net, step, x_init, sigmas = bench.make_step(128, 1, dev, 0, True)
x = x_init.clone()
with torch.no_grad():
    step(x, float(sigmas[0]), float(sigmas[1]))
    for i in range(1, 6):
        step(x, float(sigmas[i]), float(sigmas[i + 1]))
    torch.cuda.synchronize()
    K = 20
    def ev_time(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(K): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1e3 / K
    s = lambda i: (float(sigmas[5 + i % 40]), float(sigmas[6 + i % 40]))
    print("back-to-back device-resident (ms gpu, ms wall):", ev_time(lambda i: step(x, *s(i))))
    def synced(i):
        step(x, *s(i)); torch.cuda.current_stream().synchronize()
    print("device-resident + sync each step:", ev_time(synced))
    xh = x_init.cpu().pin_memory()
    print("step_host (H2D + step + D2H + sync):", ev_time(lambda i: step.step_host(xh, *s(i))))
    def replay_only(i):
        step.graph.replay(); torch.cuda.current_stream().synchronize()
    print("graph.replay + sync only:", ev_time(replay_only))
    def replay_b2b(i):
        step.graph.replay()
    print("graph.replay back-to-back:", ev_time(replay_b2b))
    t0 = time.perf_counter()
    for i in range(K): step._set_scalars(*s(i))
    print("set_scalars host ms:", (time.perf_counter() - t0) * 1e3 / K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K): step.graph.replay()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("cpu time of graph.replay() call ms:", (t1 - t0) * 1e3 / K)
