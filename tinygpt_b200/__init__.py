"""tinygpt_b200 — B200-native decode engine behind TinyGPT's generate loop.

The package is a thin host-side mirror of the reference interfaces for the decode hot path:
  tinygpt_b200.ops      ↔ tinytorch::function::{linear, rmsNorm, ropeApply, flashAttention, siluMul, add, …}
  tinygpt_b200.engine   ↔ tinygpt::GPTModel::forward / GPTEngine::generateSync (greedy)
  tinygpt_b200.models   ↔ the model constants + weight-name layout of src/model/*.h
All compute goes through the C ABI of lib/libb200decode.so (include/b200_decode.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (does not load the shared library until first use)

__all__ = ["_lib", "ops", "engine", "models", "build"]
