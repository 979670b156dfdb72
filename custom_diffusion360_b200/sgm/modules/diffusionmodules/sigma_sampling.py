"""Training-time noise-level samplers (reference: sigma_sampling.py:16-53).  Both draw an INDEX into
the discretised sigma table; `rand` injects that index tensor so that parity tests can replay the
draw the reference / oracle made."""
import torch

from ...util import instantiate_from_config


class _SigmaTable:
    """The flipped sigma table of the discretisation plus index lookup, shared by the samplers."""

    def __init__(self, discretization_config, num_idx, do_append_zero=False, flip=True):
        self.num_idx = num_idx
        discretization = instantiate_from_config(discretization_config)
        self.sigmas = discretization(num_idx, do_append_zero=do_append_zero, flip=flip)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def draw(self, n_samples):
        raise NotImplementedError

    def __call__(self, n_samples, rand=None):
        return self.idx_to_sigma(self.draw(n_samples) if rand is None else rand)


class DiscreteSampling(_SigmaTable):
    """Uniform over [num_idx_start, num_idx) (:16-34) — the reference-latent noise level (yaml: num_idx 50)."""

    def __init__(self, discretization_config, num_idx, num_idx_start=0, do_append_zero=False, flip=True):
        super().__init__(discretization_config, num_idx, do_append_zero=do_append_zero, flip=flip)
        self.num_idx_start = num_idx_start

    def draw(self, n_samples):
        return torch.randint(self.num_idx_start, self.num_idx, (n_samples,))


class CubicSampling(_SigmaTable):
    """idx = floor((1 - u^3) (num_idx - 1)), u ~ U[0, 1): biased towards high noise (:37-53)."""

    def draw(self, n_samples):
        u = torch.rand((n_samples,))
        return ((1 - u ** 3) * (self.num_idx - 1)).long()
