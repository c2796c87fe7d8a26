"""saugns_b200 -- B200-native generator back end for saugns (hot path only).

Host-side mirror of the reference generator interface (sau/generator.h:20-26)
over the C ABI in include/saugen_b200.h.  All audio is computed by the
hand-written sm_100a kernels in saugns_b200/csrc/kernels.cu; importing this
package without the built extension, or using it without a CUDA device,
raises -- there is no CPU fallback.
"""
from .generator import (Generator, LIB_PATH, lib, render, run_many, device_count,  # noqa: F401
                        WaveTables, OpView, last_error, flatten, voice_groups, default_tables)

__all__ = ["Generator", "render", "run_many", "lib", "LIB_PATH", "device_count", "WaveTables",
           "OpView", "last_error", "flatten", "voice_groups", "default_tables"]
