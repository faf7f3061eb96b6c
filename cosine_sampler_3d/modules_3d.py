"""Drop-in for the reference package `cosine_sampler_3d` (`from cosine_sampler_3d
import CosineSampler3d`, reference `cosine_sampler_3d/__init__.py:1`): re-exports the
B200-native implementation under the reference's module path."""
from cosinesampler_b200.modules_3d import (  # noqa: F401
    CosineSampler3d, CosineSamplerBackward, CosineSamplerBackwardBackward,
    padding_mode_enum, kernel_enum, _cosine_3d)
