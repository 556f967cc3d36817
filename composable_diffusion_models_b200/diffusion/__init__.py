from .samplers import SuperDiffSampler  # noqa: F401
