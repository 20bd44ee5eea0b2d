"""rayrtrace_b200 -- B200-native replacement for the `sidtrace` hot path of
AVSidorov/RAYRTRACE (the per-ray geometric-optics integrator).

    rayrtrace_b200.sidtrace   trace / check / trace_batch over the C ABI
    rayrtrace_b200.nfield     host-side field producer (den.py Nfield restated)
    rayrtrace_b200.ray        Ray adapter + batched launcher (ray.py / beam.py:340-342)
    rayrtrace_b200.beam / ft2 the callers above it (Beam, field_result, phases)
    rayrtrace_b200.dist       rays sharded over GPUs
    rayrtrace_b200.physics    extension beyond the reference: adaptive-RK rays, O/X mode, absorption
    rayrtrace_b200.dropin     directory holding a module literally named `sidtrace`

Importing the compute modules requires the CUDA library
(rayrtrace_b200/_lib/libsidtrace_b200.so); there is no CPU fallback.
"""
import os as _os

__version__ = "0.1.0"


def dropin_path() -> str:
    """Directory to put on sys.path so that ``import sidtrace`` resolves to this package."""
    return _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")
