"""Determinism probe (GPU box): the same training batch (all random draws injected) evaluated three times
without an optimiser step must give the same loss and (up to fp32 atomic ordering in two reductions) the
same gradients; also times the attention backward per shape.  Diagnostic tool."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from custom_diffusion360_b200 import ops

dev = torch.device("cuda:0")
engine, net, batch = B.make_train(64, 4, dev, 0)
opt = engine.configure_optimizers()
g = torch.Generator(device=dev).manual_seed(5)
batch["rand"] = {"sigma_idx": torch.tensor([700]), "sigma_ref_idx": torch.tensor([20]),
                 "noise": torch.randn(1, 4, 64, 64, device=dev, generator=g),
                 "noise_ref": torch.randn(1, 4, 4, 64, 64, device=dev, generator=g),
                 "noise_ref2": torch.randn(1, 4, 4, 64, 64, device=dev, generator=g)}
prev = None
for it in range(3):
    opt.zero_grad()
    loss = float(engine.training_step(dict(batch)))
    torch.cuda.synchronize()
    gr = opt.flat.grad.clone()
    print("iter", it, "loss", loss, engine.last_loss_dict, "grad norm", float(gr.norm()),
          "max |dgrad| vs prev", None if prev is None else float((gr - prev).abs().max()))
    prev = gr
# attention backward timing per shape
def t_attn(batch_, heads, nq, nkv, self_attn):
    inner = heads * 64
    q = torch.randn(batch_ * nq, inner, device=dev).bfloat16()
    kv = torch.randn(batch_ * nkv, 2 * inner, device=dev).bfloat16()
    do = torch.randn_like(q)
    k, v = kv[:, :inner], kv[:, inner:]
    o = ops.attention(q, k, v, batch_, heads, nq, nkv)
    dq = torch.empty_like(q); dkv = torch.empty_like(kv)
    kw = dict(dk=dkv[:, :inner], dv=dkv[:, inner:]) if self_attn else {}
    for _ in range(3):
        ops.attention_bwd(q, k, v, o, do, batch_, heads, nq, nkv, dq=dq, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention_bwd(q, k, v, o, do, batch_, heads, nq, nkv, dq=dq, **kw)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ops.attention_bwd(q, k, v, o, do, batch_, heads, nq, nkv, dq=dq, **kw)
    host = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    print(f"attention_bwd b{batch_} h{heads} nq{nq} nkv{nkv} self={self_attn}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us (host enqueue {host * 1e6:.0f} us)")
for shape in ((1, 10, 1024, 1024, True), (1, 20, 256, 256, True), (1, 10, 1024, 77, False), (1, 20, 256, 77, False),
              (1, 10, 24576, 77, False), (1, 20, 6144, 77, False)):
    t_attn(*shape)
