"""2D operator: same names and signatures as the reference's
`cosine_sampler_2d/modules_2d.py` (CosineSampler2d :20-44, CosineSamplerBackward
:47-74, CosineSamplerBackwardBackward :76-111, padding_mode_enum :4, kernel_enum :12).

    val = CosineSampler2d.apply(cells, grid, 'zeros', True, 'cosine', True)
"""
from . import ops as _cosine_2d  # noqa: F401  (native-surface mirror: forward/backward/...)
from .autograd import make_functions, padding_mode_enum  # noqa: F401

(CosineSampler2d, CosineSamplerBackward, CosineSamplerBackwardBackward, kernel_enum) = make_functions(2)

__all__ = ["CosineSampler2d", "CosineSamplerBackward", "CosineSamplerBackwardBackward",
           "padding_mode_enum", "kernel_enum"]
