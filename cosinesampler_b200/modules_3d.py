"""3D operator: same names and signatures as the reference's
`cosine_sampler_3d/modules_3d.py` (CosineSampler3d :20-45, CosineSamplerBackward
:47-68, CosineSamplerBackwardBackward :70-100, padding_mode_enum :4, kernel_enum :12;
the linear kernel is called 'trilinear' here).

    val = CosineSampler3d.apply(cells, grid, 'zeros', True, 'cosine', True)
"""
from . import ops as _cosine_3d  # noqa: F401
from .autograd import make_functions, padding_mode_enum  # noqa: F401

(CosineSampler3d, CosineSamplerBackward, CosineSamplerBackwardBackward, kernel_enum) = make_functions(3)

__all__ = ["CosineSampler3d", "CosineSamplerBackward", "CosineSamplerBackwardBackward",
           "padding_mode_enum", "kernel_enum"]
