"""Training-time noise-level samplers (reference: sigma_sampling.py:16-53).  `rand` injects the
random draw (index tensor) so that parity tests can replay the reference's draw."""
import torch

from ...util import default, instantiate_from_config


class DiscreteSampling:
    def __init__(self, discretization_config, num_idx, num_idx_start=0, do_append_zero=False, flip=True):
        self.num_idx = num_idx
        self.num_idx_start = num_idx_start
        self.sigmas = instantiate_from_config(discretization_config)(num_idx, do_append_zero=do_append_zero,
                                                                     flip=flip)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def __call__(self, n_samples, rand=None):
        idx = default(rand, lambda: torch.randint(self.num_idx_start, self.num_idx, (n_samples,)))
        return self.idx_to_sigma(idx)


class CubicSampling:
    """idx = floor((1 - u^3) (num_idx - 1)), u ~ U[0, 1): biased towards high noise (:37-53)."""

    def __init__(self, discretization_config, num_idx, do_append_zero=False, flip=True):
        self.num_idx = num_idx
        self.sigmas = instantiate_from_config(discretization_config)(num_idx, do_append_zero=do_append_zero,
                                                                     flip=flip)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def __call__(self, n_samples, rand=None):
        if rand is None:
            t = torch.rand((n_samples,))
            rand = ((1 - t ** 3) * (self.num_idx - 1)).long()
        return self.idx_to_sigma(rand)
