"""Classifier-free guidance batch assembly / combination (reference: guiders.py:102-166).

`rows` is the number of UNet rows per image; row order is the reference's:
ScheduledCFGImgTextRef -> (uncond, image-cond-only, full cond), VanillaCFGImgRef -> (uncond, cond).
At inference the *_ref halves of the conditioning are empty, so only the first x.size(0) rows of
every cond tensor are used (the `split` in the reference yields an empty second part)."""
import torch


class _CFGBase:
    rows = 1

    def _assemble(self, x, s, c, uc, order):
        c_out = {}
        nb = x.size(0)
        for k in c:
            if k in ("vector", "crossattn", "concat"):
                uc1, uc2 = uc[k][:nb], uc[k][nb:]
                c1, c2 = c[k][:nb], c[k][nb:]
                first = {"u": uc1, "c": c1}
                second = {"u": uc2, "c": c2}
                # target rows for every guidance row, then the (possibly empty) reference rows
                parts = [first[o] for o in order] + [second[o] for o in order[:1]] + \
                        [second["c"] for _ in order[1:]]
                c_out[k] = torch.cat(parts, 0)
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * len(order)), torch.cat([s] * len(order)), c_out


class ScheduledCFGImgTextRef(_CFGBase):
    """InstructPix2Pix-style two-scale guidance: x_u + s (x_c - x_ic) + s_im (x_ic - x_u)."""
    rows = 3

    def __init__(self, scale: float, scale_im: float):
        self.scale = scale
        self.scale_im = scale_im

    def prepare_inputs(self, x, s, c, uc):
        return self._assemble(x, s, c, uc, "uuc")

    def __call__(self, x, sigma):
        x_u, x_ic, x_c = x.chunk(3)
        return x_u + self.scale * (x_c - x_ic) + self.scale_im * (x_ic - x_u)


class VanillaCFGImgRef(_CFGBase):
    rows = 2

    def __init__(self, scale: float):
        self.scale = scale
        self.scale_im = 0.0

    def prepare_inputs(self, x, s, c, uc):
        return self._assemble(x, s, c, uc, "uc")

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        return x_u + self.scale * (x_c - x_u)


class IdentityGuider:
    rows = 1
    scale = 1.0
    scale_im = 0.0

    def prepare_inputs(self, x, s, c, uc):
        return x, s, {k: c[k] for k in c}

    def __call__(self, x, sigma):
        return x
