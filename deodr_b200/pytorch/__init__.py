"""PyTorch surface of the rasteriser (mirror of ``deodr.pytorch`` for the hot path)."""
from .differentiable_renderer_pytorch import (  # noqa: F401
    CudaDifferentiableRender2D,
    CudaDifferentiableRenderer2DFunc,
    TorchDifferentiableRender2D,
    TorchDifferentiableRenderer2DFunc,
)
