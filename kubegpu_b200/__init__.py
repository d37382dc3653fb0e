"""kubegpu_b200 -- B200-native topology-aware GPU placement scorer.

Drop-in for the scoring path of microsoft/KubeGPU's ``gpuschedulerplugin``
(``DeviceScheduler.PodFitsDevice`` / ``PodAllocate``, see include/kgpu.h): hand-written
sm_100a CUDA kernels behind a C ABI (``libkgpu.so``).  This package is only the thin
Python face used by tests and bench.py: ``_lib`` (ctypes binding of the C ABI),
``scorer`` (handle wrapper; host and torch-device entry points), ``synth`` (seeded
synthetic inputs).  There is NO CPU fallback: if ``libkgpu.so`` is missing the import
of ``_lib`` raises, and without a CUDA device ``Scorer()`` raises.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
__version__ = "0.1.0"
