"""DiffusionEngine surface (reference: sgm/models/diffusion.py:43-557) around the B200 UNet.

Kept: constructor config keys, `.model` (OpenAIWrapper) / `.model.diffusion_model`, `.denoiser`,
`.sampler`, `sample(cond, uc, batch_size, num_steps, randn, shape, **kwargs)`,
`clear_rendered_feat()`, trainable-parameter selection by name (`trainkeys`).
The text conditioner (sgm/modules/encoders/modules.py) and the decode-only first stage are built
when their `target:` names this package; otherwise callers feed `cond` / `uc` embeddings and receive
latents.  Lightning is not required (plain nn.Module).

Training (`training_step`, reference :221-272): there is no autograd on this path — `forward`
evaluates the loss through the taped UNet forward and immediately runs the explicit backward
(sgm/modules/train_path.py), leaving the gradients of the trainable (pose) parameters in `.grad`;
`configure_optimizers()` returns the fused AdamW over the flattened pose parameters
(sgm/optim.py), whose `step()` also performs the data-parallel gradient all-reduce.  Inputs are
LATENTS (the VAE is outside SURVEY §8): batch[input_key] = x [b,4,L,L], batch[input_key + "_ref"]
= reference latents [b,n,4,L,L], batch["cond"] = {"crossattn": [b+b*n,77,ctx], "vector":
[b+b*n,adm]} (what the conditioner would emit), plus pose / mask / depth (opacity) / drop_im /
rgb (the image for the rgb term) as in the reference's data loader.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from ..modules.diffusionmodules.sampling import FusedGuidedStep
from ..modules.diffusionmodules.wrappers import OPENAIUNETWRAPPER
from ..util import default, get_obj_from_str, instantiate_from_config
from .autoencoder import AutoencoderKL as _OwnAutoencoder


def _cfg_get(cfg, path, dflt=None):
    cur = cfg
    for key in path:
        try:
            cur = cur[key]
        except (KeyError, TypeError, IndexError):
            return dflt
    return cur


class DiffusionEngine(nn.Module):
    def __init__(self, network_config, denoiser_config, first_stage_config=None,
                 conditioner_config=None, sampler_config=None, optimizer_config=None,
                 scheduler_config=None, loss_fn_config=None, network_wrapper=None, ckpt_path=None,
                 use_ema=False, ema_decay_rate=0.9999, scale_factor=1.0,
                 disable_first_stage_autocast=False, input_key="jpg", log_keys=None,
                 no_cond_log=False, compile_model=False, trainkeys="pose", multiplier=0.05,
                 loss_rgb_lambda=20.0, loss_fg_lambda=10.0, loss_bg_lambda=20.0):
        super().__init__()
        if use_ema:
            raise NotImplementedError("use_ema=True is unused by the shipped config")
        if trainkeys != "pose":
            # 'poseattn' / 'all' would hand frozen SDXL weights to AdamW with zero gradients (the explicit
            # backward only produces pose + conditioning gradients): decoupled weight decay would then
            # shrink them every step.  The shipped config trains `pose` (train_co3d_concept.yaml:8).
            raise NotImplementedError(f"trainkeys={trainkeys!r}: only 'pose' (the shipped config) is built")
        if ckpt_path is not None:
            raise NotImplementedError("ckpt_path: load checkpoints with sgm.util.load_checkpoints(engine, base_sd, delta_sd)")
        if scheduler_config is not None:
            raise NotImplementedError("scheduler_config is None in the shipped config; LR schedules are not built")
        self.log_keys = log_keys
        self.input_key = input_key
        self.trainkeys = trainkeys
        self.multiplier = multiplier
        self.loss_rgb_lambda, self.loss_fg_lambda, self.loss_bg_lambda = loss_rgb_lambda, loss_fg_lambda, loss_bg_lambda
        self.rgb = _cfg_get(network_config, ("params", "rgb"), False)
        self.rgb_predict = _cfg_get(network_config, ("params", "rgb_predict"), False)
        self.scale_factor = scale_factor
        self.use_ema = False
        model = instantiate_from_config(network_config)
        self.model = get_obj_from_str(default(network_wrapper, OPENAIUNETWRAPPER))(model, compile_model=compile_model)
        self.denoiser = instantiate_from_config(denoiser_config)
        self.sampler = instantiate_from_config(sampler_config) if sampler_config is not None else None
        # deliberately not built (out of the hot path): kept as configs for a caller that wants them
        self.conditioner_config = conditioner_config
        self.first_stage_config = first_stage_config
        self.loss_fn_config = loss_fn_config
        self.loss_fn = instantiate_from_config(loss_fn_config) if loss_fn_config is not None else None
        self.optimizer_config = optimizer_config if optimizer_config is not None else {"target": "torch.optim.AdamW"}
        self.scheduler_config = scheduler_config
        self.learning_rate = 1.0e-4   # base_learning_rate of the shipped yaml; main.py:1019-1050 overrides
        self.global_step = 0
        # the conditioner of THIS package (text towers on the CUDA kernels) is built when
        # its target names it; with the reference's own `sgm.modules.GeneralConditioner` target the caller
        # keeps the reference object and assigns it to `.conditioner`
        self.conditioner = None
        if conditioner_config is not None and str(conditioner_config.get("target", "")).startswith("custom_diffusion360_b200."):
            self.conditioner = instantiate_from_config(conditioner_config)
        self.first_stage_model = None
        self.last_cond_grads = None
        # trainable set by parameter name (reference :119-147)
        for name, p in self.model.diffusion_model.named_parameters():
            p.requires_grad = "pose" in name
        self._fused: Optional[FusedGuidedStep] = None

    @property
    def device(self):
        return next(self.model.parameters()).device

    def clear_rendered_feat(self):
        self.model.diffusion_model.clear_rendered_feat()

    def set_reference_choices(self, choices):
        self.model.diffusion_model.set_reference_choices(choices)

    def init_first_stage(self, config=None):
        """Instantiate the first stage for `decode_first_stage` (reference :191-202).  Only targets of
        this package are built (the decode-only AutoencoderKL of sgm/models/autoencoder.py); with the
        reference's own `sgm.models.autoencoder.*` target the caller keeps the reference object and
        assigns it to `.first_stage_model` itself."""
        config = default(config, self.first_stage_config)
        if config is None or not str(config.get("target", "")).startswith("custom_diffusion360_b200."):
            raise NotImplementedError("first_stage_config.target must be custom_diffusion360_b200.sgm.models."
                                      "autoencoder.AutoencoderKLInferenceWrapper to be built here")
        model = instantiate_from_config(config).eval()
        for p_ in model.parameters():
            p_.requires_grad = False
        self.first_stage_model = model.to(self.device)
        return self.first_stage_model

    @torch.no_grad()
    def decode_first_stage(self, z):
        """z / scale_factor -> first_stage_model.decode (reference :207-212); the division is folded
        into the 1x1 post_quant_conv of the decode kernel chain."""
        if self.first_stage_model is None:
            self.init_first_stage()
        fs = self.first_stage_model
        if isinstance(fs, _OwnAutoencoder):
            return fs.decode(z, scale=1.0 / self.scale_factor)
        return fs.decode(1.0 / self.scale_factor * z)      # a reference first stage assigned by the caller

    @torch.no_grad()
    def sample(self, cond: Dict, uc: Union[Dict, None] = None, batch_size: int = 16, num_steps=None,
               randn=None, shape: Union[None, Tuple, List] = None, return_rgb=False, mask=None,
               init_im=None, noise=None, fused: bool = True, **kwargs):
        """Reference contract (diffusion.py:375-401).  `noise` is accepted as an alias of `randn`
        (sample.py:190-192 passes `noise=`).  With `fused=True` (default) the loop runs the
        graph-replayed fused step; `fused=False` goes through the generic denoiser/sampler objects."""
        if mask is not None or init_im is not None:
            raise NotImplementedError("masked / img2img sampling is unused by sample.py")
        randn = noise if randn is None else randn
        if randn is None:
            randn = torch.randn(batch_size, *shape)
        x = randn.to(self.device).float().contiguous().clone()
        uc = default(uc, cond)
        if not fused:
            denoiser = lambda inp, sigma, c: self.denoiser(self.model, inp, sigma, c, **kwargs)
            samples, rgb_list = self.sampler(denoiser, x, cond, uc=uc, num_steps=num_steps)
            return (samples, rgb_list) if return_rgb else samples
        n_img = x.shape[0]
        pose = kwargs.get("pose")
        if isinstance(pose, (list, tuple)):
            pose = list(pose)[:n_img]  # sample.py passes `pose * rows`
        # Images of the same shape re-use ONE FusedGuidedStep: its conditioning / camera buffers are
        # overwritten in place and both captured graphs (step 0 with FeatureNeRF, steady state) are replayed
        # — a fresh object per call would pay an eager step 0, an eager steady step and a capture (~150 ms)
        # on every image of a sweep (sample.py calls sample() once per target pose).
        step = self._fused
        if (step is not None and pose is not None
                and step.matches(self.model.diffusion_model, self.sampler.guider, cond, uc, pose, n_img, tuple(x.shape[1:]))):
            step.set_cond(cond, uc)
            step.set_pose(pose)
        else:
            step = FusedGuidedStep(self.model.diffusion_model, self.denoiser, self.sampler.guider, cond, uc,
                                   pose=pose, n_img=n_img, latent_shape=tuple(x.shape[1:]))
            self._fused = step
        samples = self.sampler.sample_fused(step, x, num_steps=num_steps)
        if step.x_static is not None and samples.data_ptr() == step.x_static.data_ptr():
            samples = samples.clone()     # the step object keeps that buffer for the next image's replays
        return (samples, None) if return_rgb else samples

    # ---- training (reference :204-272, 310-373) ----------------------------------------------------
    def get_input(self, batch):
        k = self.input_key
        return (batch[k], batch.get(k + "_ref"), batch.get("pose"), batch.get("mask"), batch.get("mask_ref"),
                batch.get("depth"), batch.get("drop_im", 0.0))

    def forward(self, x, x_rgb, xr, pose, mask, mask_ref, opacity, drop_im, batch, backward: bool = True,
                sync: bool = True):
        """Loss of one batch (reference :221-236) and — the autograd replacement — the gradients of
        the trainable parameters in `.grad`.  Returns (loss_mean tensor, loss_dict).  Everything is
        evaluated on the device (the reference's `if loss_rgb.mean() > 0` becomes a 0/1 factor), so
        with sync=False nothing waits for the GPU and the call can be captured in a CUDA graph;
        loss_dict then holds device scalars instead of floats."""
        loss, loss_fg, loss_bg, loss_rgb = self.loss_fn(self.model, self.denoiser, self.conditioner, x, x_rgb,
                                                        xr, pose, mask, mask_ref, opacity, batch)
        b = x.shape[0]
        dev = x.device
        loss_mean = loss.mean()
        terms = {"loss": loss_mean}
        if torch.is_tensor(drop_im):
            drop = drop_im.float().reshape(-1).to(dev).expand(b)
        else:
            drop = torch.full((b,), float(drop_im), device=dev)
        dsum = drop.sum() + 1e-12
        w_fg = w_bg = w_rgb = None
        if self.rgb and self.global_step > 0 and torch.is_tensor(loss_fg):
            k = loss_fg.shape[1]
            lf = (loss_fg.mean(1) * drop).sum() / dsum
            lb = (loss_bg.mean(1) * drop).sum() / dsum
            loss_mean = loss_mean + self.loss_fg_lambda * lf + self.loss_bg_lambda * lb
            terms["loss_fg"], terms["loss_bg"] = lf, lb
            w_fg = self.loss_fg_lambda * drop / (k * dsum)
            w_bg = self.loss_bg_lambda * drop / (k * dsum)
        if self.rgb_predict and torch.is_tensor(loss_rgb):
            k = loss_rgb.shape[1]
            gate = (loss_rgb.mean() > 0).float()                      # reference :232
            lr_ = (loss_rgb.mean(1) * drop).sum() / dsum
            loss_mean = loss_mean + gate * self.loss_rgb_lambda * lr_
            terms["loss_rgb"] = lr_
            w_rgb = gate * self.loss_rgb_lambda * drop / (k * dsum)
        if backward:
            self._route_cond_grads(self.loss_fn.backward(1.0 / b, w_fg, w_bg, w_rgb), self.loss_fn.last_cond)
        if sync:
            terms = {k_: float(v) for k_, v in terms.items()}
        return loss_mean, terms

    def _route_cond_grads(self, grads, cond):
        """Conditioning gradients of the step: `self.last_cond_grads` = {"crossattn": [b + b*n, 77, ctx],
        "vector": [b + b*n, adm]} fp32, shaped like the conditioner's outputs (rows of the reference
        views are zero: that stream is no_grad in the reference).  When those outputs came from a
        torch-autograd conditioner — the reference's GeneralConditioner with its trainable `<new1>` token
        rows (reference :343-356, main.py:627-643) — the gradients are pushed into that graph, which is
        what the reference's `loss.backward()` does for the text encoders."""
        self.last_cond_grads = None
        if grads is None:
            return
        full = {}
        for k, g in grads.items():
            ref = cond[k]
            out = torch.zeros(ref.shape, device=g.device, dtype=torch.float32)
            out[: g.shape[0]] = g.reshape(g.shape[0], *ref.shape[1:])
            full[k] = out
        self.last_cond_grads = full
        live = [(cond[k], full[k]) for k in full if cond[k].requires_grad]
        if live and not torch.cuda.is_current_stream_capturing():
            torch.autograd.backward([t for t, _ in live], [g.to(t.dtype) for t, g in live])

    def shared_step(self, batch, backward: bool = True, sync: bool = True):
        x, xr, pose, mask, mask_ref, opacity, drop_im = self.get_input(batch)
        x_rgb = batch.get("rgb")
        if xr is not None and torch.is_tensor(drop_im):
            bb = xr.shape[0]
            xr = drop_im.reshape(bb, 1, 1, 1, 1).to(xr) * xr          # reference :246
        batch["global_step"] = self.global_step
        return self(x, x_rgb, xr, pose, mask, mask_ref, opacity, drop_im, batch, backward=backward, sync=sync)

    def training_step(self, batch, batch_idx=0):
        """Loss + gradients of one batch (the Lightning hook of the reference, :251-272; here the
        caller owns the loop: `opt.zero_grad(); loss = engine.training_step(batch); opt.step()`)."""
        loss, loss_dict = self.shared_step(batch)
        self.last_loss_dict = loss_dict
        return loss

    def configure_optimizers(self, group=None):
        """AdamW over the trainable parameters selected by `trainkeys` (reference :310-373; the
        conditioner's token rows are outside this build).  `group`: torch.distributed process group
        for the data-parallel gradient all-reduce (None = default group if initialised)."""
        from ..optim import PoseAdamW
        cfg = dict(self.optimizer_config.get("params", {}))
        target = self.optimizer_config.get("target", "torch.optim.AdamW")
        if not target.endswith("AdamW"):
            raise NotImplementedError(f"optimizer {target}: only AdamW (the reference default) is built")
        named = [(n, p) for n, p in self.model.diffusion_model.named_parameters() if p.requires_grad]
        opt = PoseAdamW(named, lr=self.learning_rate, group=group, **cfg)
        opt.on_step = self._after_optimizer_step
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            opt.attach_overlap(self.model.diffusion_model)
        # the parameters were re-homed into the optimiser's flat buffer: packs built before hold views
        # of / copies from the old storage
        from ..modules.attention import invalidate_all_packed
        invalidate_all_packed(self.model.diffusion_model, only_trainable=True, keep_buffers=False)
        return opt

    def _after_optimizer_step(self):
        self.global_step += 1
        from ..modules.attention import invalidate_all_packed
        invalidate_all_packed(self.model.diffusion_model, only_trainable=True)


class GraphedTrainStep:
    """The training step (noising -> reference stream -> taped forward -> losses -> backward to the
    pose gradients) captured ONCE in a CUDA graph and replayed per batch: the eager step issues
    ~3.6 k launches and is bound by the host (~26 us per launch from Python); replay is bound by
    the GPU.  Per call: the batch and the step's random draws (sigma indices from the reference's
    samplers, three noise tensors, stratified ray / depth variates) are written into static
    buffers, the graph replays, then the optimiser runs (gradient all-reduce + fused AdamW).
    The bf16 operand packs of the TRAINABLE weights live in persistent buffers that are refreshed in
    place from the fp32 masters inside the graph (attention._Packed.mark_stale), so every replay sees
    the previous optimiser update and the eager step shares the same buffers; the nviews bias is read
    from device memory for the same reason."""

    def __init__(self, engine: "DiffusionEngine", opt, batch: dict, comm: str = "graph"):
        """comm (data-parallel runs only): "graph" — the per-pose-block bucket all-reduces (NCCL) are
        CAPTURED inside the step's graph on the communication stream, each forked off the backward walk
        the moment that block's gradients are written, so the exchange overlaps the rest of the backward
        and a replay ends with reduced gradients; "after" — no collective inside the graph, all buckets
        are reduced after the replay (round-1 behaviour); "none" — no exchange at all (measurement of
        the exposed communication time only)."""
        from ..modules.attention import invalidate_all_packed
        from ..modules.utils_cameraray import pack_pose
        self.engine, self.opt = engine, opt
        unet = engine.model.diffusion_model
        dev = engine.device
        k = engine.input_key
        self.keys = [key for key in (k, k + "_ref", "mask", "mask_ref", "depth", "rgb", "drop_im") if torch.is_tensor(batch.get(key))]
        self.static = {key: batch[key].to(dev).clone() for key in self.keys}
        self.static["pose"] = pack_pose(batch["pose"], dev).clone()
        self.static_cond = {n: t.to(dev).clone() for n, t in batch["cond"].items()}
        x, xr = self.static[k], self.static[k + "_ref"]
        b = x.shape[0]
        self.rand = dict(sigma=torch.ones(b, device=dev), sigma_ref=torch.ones(b, device=dev),
                         noise=torch.zeros_like(x), noise_ref=torch.zeros_like(xr), noise_ref2=torch.zeros_like(xr))
        self.stratified = []
        for _, blk in unet.pose_blocks():
            rm = blk.pose_featurenerf.raymarcher
            if not rm.stratified:
                self.stratified = None
                break
            c = blk.pose_emb_layers.weight.shape[0]
            res = x.shape[-1] // (c // unet.model_channels)
            hw, d = res * res, rm.num_samples
            self.stratified.append(dict(res=res, rm=rm, bins=(torch.zeros(hw, 2, device=dev), torch.zeros(hw, d, device=dev),
                                                              torch.zeros(hw, d, device=dev))))
        if self.stratified:
            self.rand["jitter"] = [{"bins": s_["bins"]} for s_ in self.stratified]
        if engine.global_step < 1:
            raise RuntimeError("run the first optimisation step eagerly: at global_step 0 the reference "
                               "leaves the fg / bg terms out of the total (diffusion.py:225)")
        self._draw()
        self.opt.zero_grad()
        assert comm in ("graph", "after", "none")
        world = self.opt._world()
        self.comm = comm if world > 1 else "none"
        self.comm_in_graph = self.comm == "graph"
        # "after" / "none": no collectives inside the captured step (the per-block callbacks of the
        # backward walk do nothing); "graph": they fire during capture and are recorded
        self.opt.suspend_overlap = not self.comm_in_graph
        engine.shared_step(self._batch(), sync=False)              # warm-up: builds every frozen pack
        self.opt.wait_reduce()
        invalidate_all_packed(unet, only_trainable=True)           # ... the trainable ones are refreshed IN PLACE inside the graph
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.terms = engine.shared_step(self._batch(), sync=False)
            self.opt.wait_reduce()                                 # join the communication stream into the capture
        # conditioning gradients of the replayed step (static buffers of the graph's pool, rewritten by
        # every replay): {"crossattn", "vector"} shaped like batch["cond"], or None when switched off
        self.cond_grads = engine.last_cond_grads
        self.opt.suspend_overlap = True                            # nothing eager may start a second exchange

    def _batch(self):
        bt = dict(self.static)
        bt["cond"] = self.static_cond
        bt["rand"] = self.rand
        return bt

    @torch.no_grad()
    def _draw(self):
        """The step's random draws, with the reference's samplers / formulas, into the static buffers."""
        lf = self.engine.loss_fn
        b = self.rand["sigma"].shape[0]
        self.rand["sigma"].copy_(lf.sigma_sampler(b))
        self.rand["sigma_ref"].copy_(lf.sigma_sampler_ref(b))
        for n_ in ("noise", "noise_ref", "noise_ref2"):
            self.rand[n_].normal_()
        for s_ in self.stratified or []:
            res, rm = s_["res"], s_["rm"]
            xy, depths, dists = s_["bins"]
            dev = xy.device

            def positions():   # get_patch_raybundle, stratified (utils_cameraray.py:111-140)
                edges = torch.linspace(1, -1, res + 1, device=dev)
                center = (edges[1:] + edges[:-1]) / 2.0
                upper = torch.cat([center, edges[-1:]], -1)
                lower = torch.cat([edges[:1], center], -1)
                return (lower + (upper - lower) * torch.rand(res + 1, device=dev))[:-1]

            hpos, vpos = positions(), positions()
            xy[:, 0].copy_(hpos[None, :].expand(res, res).reshape(-1))
            xy[:, 1].copy_(vpos[:, None].expand(res, res).reshape(-1))
            lower = rm.lengths_lower.to(device=dev, dtype=torch.float32)
            upper = rm.lengths_upper.to(device=dev, dtype=torch.float32)
            jit = lower[None] + (upper - lower)[None] * torch.rand(depths.shape[0], depths.shape[1] + 1, device=dev)
            depths.copy_((jit[:, :-1] + jit[:, 1:]) / 2.0)            # Raymarcher.stratified_sampling (:317-325)
            dists.copy_(jit[:, 1:] - jit[:, :-1])

    @torch.no_grad()
    def __call__(self, batch: dict, step_optimizer: bool = True):
        """One optimisation step on `batch`; returns the (device) total loss of the step."""
        from ..modules.utils_cameraray import pack_pose
        for key in self.keys:
            self.static[key].copy_(batch[key], non_blocking=True)
        self.static["pose"].copy_(pack_pose(batch["pose"], self.static["pose"].device), non_blocking=True)
        for n_, t in batch["cond"].items():
            self.static_cond[n_].copy_(t, non_blocking=True)
        self._draw()
        self.graph.replay()
        if step_optimizer:
            self.opt.step(reduce=self.comm == "after")
        return self.loss
