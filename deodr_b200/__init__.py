"""deodr_b200 - B200 (sm_100a) replacement of DEODR's raster core ``renderScene`` / ``renderScene_B``.

Layout (hot path only):

* ``csrc/``                              hand-written CUDA kernels + the C-ABI of ``include/deodr_b200.h``
* ``differentiable_renderer_cython``     drop-in for the reference FFI module (``renderSceneCpp`` / ``renderSceneBCpp``)
* ``differentiable_renderer``            ``Scene2DBase`` / ``Scene2D`` / ``renderScene`` / ``renderSceneB`` mirrors
* ``pytorch``                            autograd Functions (reference contract + zero-copy CUDA variant)
* ``renderer``                           device-resident front-end (``DeviceScene``, ``Renderer``)
* ``distributed``                        view-sharded multi-GPU rendering with one NCCL all-reduce of shared gradients
* ``scenes``                             seeded synthetic scenes for tests and the benchmark

Importing the package does not import torch or touch CUDA; the first render does, and fails loudly without a GPU.
"""

__version__ = "0.1.0"
