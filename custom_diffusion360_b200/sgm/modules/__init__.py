"""`sgm.modules` namespace of the reference: `GeneralConditioner` is addressed as
`sgm.modules.GeneralConditioner` by the shipped yaml (configs/train_co3d_concept.yaml:57)."""


def __getattr__(name):
    if name == "GeneralConditioner":
        from .encoders.modules import GeneralConditioner
        return GeneralConditioner
    raise AttributeError(name)
