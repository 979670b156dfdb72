"""First-stage (VAE) decode on the CUDA path against the CPU oracle (oracle/vae_oracle.py, pinned
against the reference's own Decoder) and the committed golden vector; per-kernel checks of the two
new kernels and of the implicit convolution on images wider than 128 pixels.

Tolerance: bf16 activations through ~30 sequential convolutions / norms: rel_rms <= 3e-2 and
max_abs <= 0.15 max|ref| (the UNet's per-tensor bound); single kernels <= 1.5 * 2^-9."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as V

gpu = pytest.mark.gpu
BF16_EPS = 2.0 ** -9
METRICS = {}


def _record(name, **kw):
    METRICS[name] = kw
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "vae_parity_metrics.json"), "w") as f:
        json.dump(METRICS, f, indent=1)


def _rel_rms(a, b):
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


def _vae(cfg, sd, dev):
    from custom_diffusion360_b200.sgm.models.autoencoder import AutoencoderKLInferenceWrapper
    vae = AutoencoderKLInferenceWrapper(embed_dim=4, ddconfig=cfg, lossconfig={"target": "torch.nn.Identity"}).eval()
    missing, _ = vae.load_decode_state_dict(sd)
    assert not missing
    return vae.to(dev)


@gpu
def test_softmax_rows_kernel():
    from custom_diffusion360_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    for rows, n, scale in ((5, 256, 0.0625), (64, 16384, 512 ** -0.5), (3, 1000, 1.0)):
        s = 20 * torch.randn(rows, n, device=dev, generator=g)
        p = ops.softmax_rows(s, scale=scale).float()
        ref = torch.softmax(s * scale, -1)
        assert float((p - ref).abs().max()) <= 1.5 * BF16_EPS * float(ref.max()) + 1e-7
        assert float((p.sum(-1) - 1).abs().max()) <= 4e-3
    s = torch.full((2, 64), -1e30, device=dev)
    s[:, 3] = 0.0
    p = ops.softmax_rows(s).float()
    assert float(p[:, 3].min()) == 1.0 and float(p.sum()) == 2.0


@gpu
def test_pointwise_conv_kernel():
    from custom_diffusion360_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(3, 4, 17, 9, device=dev, generator=g)
    w = torch.randn(4, 4, device=dev, generator=g)
    b = torch.randn(4, device=dev, generator=g)
    out = ops.pointwise_conv_nchw(x, w, b, scale=7.677)
    ref = F.conv2d(x * 7.677, w[:, :, None, None], b)
    assert out.shape == ref.shape and float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@gpu
def test_conv3x3_on_images_wider_than_128():
    """Implicit-GEMM convolution with 128-pixel row segments (W = 256, 512): borders come from TMA
    zero fill at x = -1 / x = W only, interior segment joins read the neighbouring pixels."""
    from custom_diffusion360_b200 import ops
    from custom_diffusion360_b200.sgm.prepack import pack_conv3x3
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2)
    for (b, c, h, w, n) in ((1, 64, 8, 256, 64), (2, 128, 4, 512, 192)):
        x = torch.randn(b, c, h, w, device=dev, generator=g)
        wt = torch.randn(n, c, 3, 3, device=dev, generator=g) / (3 * c ** 0.5)
        bias = torch.randn(n, device=dev, generator=g)
        xt = x.permute(0, 2, 3, 1).reshape(-1, c).to(torch.bfloat16).contiguous()
        out = ops.conv3x3(xt, pack_conv3x3(wt), b, h, w, bias=bias).float().view(b, h, w, n).permute(0, 3, 1, 2)
        ref = F.conv2d(xt.float().view(b, h, w, c).permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), bias, padding=1)
        assert float((out - ref).abs().max()) <= 1.5 * BF16_EPS * float(ref.abs().max()) + 1e-3


@gpu
def test_vae_decode_vs_oracle_and_golden():
    dev = torch.device("cuda:0")
    from tests.golden.make_vae_golden import latent
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vae_decoder_golden.pt"), weights_only=False)
    cfg = dict(V.TINY_VAE_CFG)
    sd = V.synthetic_state_dict(cfg, seed=g["seed_w"])
    z = latent(g["seed_z"], g["batch"], g["latent"])
    vae = _vae(cfg, sd, dev)
    img = vae.decode(z.to(dev), scale=1.0 / g["scale_factor"]).float().cpu()
    ref = V.decode_first_stage(sd, cfg, z, g["scale_factor"])
    gold = g["image"].float()
    rr, rg = _rel_rms(img, ref), _rel_rms(img, gold)
    ma = float((img - ref).abs().max() / ref.abs().max())
    _record("tiny_decode", rel_rms_vs_oracle=rr, rel_rms_vs_reference_golden=rg, max_abs_over_max=ma)
    assert img.shape == ref.shape
    assert rr <= 3e-2 and rg <= 3e-2 and ma <= 0.15, (rr, rg, ma)


@gpu
def test_engine_decode_first_stage_sdxl_size():
    """The shipped first-stage config at full size: a 128x128 latent -> 1024x1024 image through
    DiffusionEngine.decode_first_stage, against the oracle functions executed in fp32 by torch on the
    same GPU (the CPU oracle needs ~10 TFLOP for this size); same tolerance."""
    from tests.test_train_step_gpu import _engine
    from oracle import sgm_oracle as O
    dev = torch.device("cuda:0")
    cfg = dict(V.SDXL_VAE_CFG)
    sd = V.synthetic_state_dict(cfg, seed=5)
    engine = _engine(dict(O.TINY_CFG), O.synthetic_state_dict(dict(O.TINY_CFG), seed=3), dev)
    engine.scale_factor = V.SDXL_SCALE_FACTOR
    engine.first_stage_config = {"target": "custom_diffusion360_b200.sgm.models.autoencoder.AutoencoderKLInferenceWrapper",
                                 "params": {"embed_dim": 4, "monitor": "val/rec_loss", "ddconfig": cfg,
                                            "lossconfig": {"target": "torch.nn.Identity"}}}
    fs = engine.init_first_stage()
    fs.load_decode_state_dict(sd)
    z = 0.13025 * torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(11)).to(dev)
    img = engine.decode_first_stage(z)
    torch.cuda.synchronize()
    assert img.shape == (1, 3, 1024, 1024) and bool(torch.isfinite(img).all())
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = V.decode_first_stage({k: v.to(dev) for k, v in sd.items()}, cfg, z, V.SDXL_SCALE_FACTOR)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    rr = _rel_rms(img.float(), ref)
    ma = float((img.float() - ref).abs().max() / ref.abs().max())
    _record("sdxl_decode_1024", rel_rms_vs_fp32=rr, max_abs_over_max=ma)
    assert rr <= 3e-2 and ma <= 0.15, (rr, ma)
