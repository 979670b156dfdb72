"""Loss weights λ(σ) (reference: denoiser_weighting.py:22-24)."""


class EpsWeighting:
    def __call__(self, sigma):
        return sigma ** -2.0
