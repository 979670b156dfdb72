"""Network wrapper: dict conditioning -> UNet kwargs (reference: wrappers.py:23-34)."""
import torch
import torch.nn as nn

OPENAIUNETWRAPPER = "custom_diffusion360_b200.sgm.modules.diffusionmodules.wrappers.OpenAIWrapper"


class IdentityWrapper(nn.Module):
    def __init__(self, diffusion_model, compile_model: bool = False):
        super().__init__()
        if compile_model:
            raise NotImplementedError("torch.compile is not used: the step is captured in a CUDA graph instead")
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


class OpenAIWrapper(IdentityWrapper):
    def forward(self, x: torch.Tensor, t: torch.Tensor, c: dict, **kwargs):
        concat = c.get("concat")
        if concat is not None and concat.numel() > 0:
            x = torch.cat((x, concat.type_as(x)), dim=1)
        return self.diffusion_model(x, timesteps=t, context=c.get("crossattn", None),
                                    y=c.get("vector", None), **kwargs)

    def train_forward(self, x: torch.Tensor, t: torch.Tensor, c: dict, **kwargs):
        """Taped forward of the training step (UNetModel.forward_train); same dict -> kwargs mapping."""
        return self.diffusion_model.forward_train(x, timesteps=t, context=c.get("crossattn", None),
                                                  y=c.get("vector", None), **kwargs)
