"""Drop-in for the reference package `cosine_sampler_2d` (`from cosine_sampler_2d
import CosineSampler2d`, reference `cosine_sampler_2d/__init__.py:1`): re-exports the
B200-native implementation under the reference's module path."""
from cosinesampler_b200.modules_2d import (  # noqa: F401
    CosineSampler2d, CosineSamplerBackward, CosineSamplerBackwardBackward,
    padding_mode_enum, kernel_enum, _cosine_2d)
