"""sst_core_b200 -- a B200-native conservative parallel discrete-event engine behind SST-core's plugin surface.

Only the event-delivery hot path is implemented: the `sst` model-description module subset
(config.py), wire-up to flat arrays (wireup.py, models.py), and the CUDA engine reached through the C-ABI in
include/sstb200.h (engine.py).  There is no CPU fallback: importing `sst_core_b200.engine` without the built
libsstb200.so raises.
"""
__all__ = ["config", "elements", "wireup", "models", "timelord"]
