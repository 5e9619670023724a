"""mygpuraytracer_b200 -- B200-native wavefront path-tracing loop.

A from-scratch sm_100a implementation of the hot path of nkkk98/MyGPURaytracer
(``pathtraceInit`` / ``pathtrace`` / ``pathtraceFree``, apps/src/pathtrace.h:6-10)
behind a C ABI (``include/b2pt.h``).  This package holds the CUDA sources
(``csrc/``), their build driver and a thin ctypes host mirror of the reference
interface (:mod:`mygpuraytracer_b200.api`).  There is no CPU fallback: using
the API without the compiled ``libb2pt.so`` raises.
"""
__version__ = "0.1.0"
