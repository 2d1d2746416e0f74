"""xagents_b200 -- B200-native on-policy rollout-to-update hot path of abstractguy/xagents.

Only what that path needs lives here:

* `csrc/`    hand-written sm_100a CUDA kernels behind the C ABI of `include/xagents_b200.h`
* `_ffi`     ctypes binding of that ABI;  `_dlpack`  zero-copy device pointers from DLPack capsules
* `ops`      tensor-level calls (returns/GAE, permute-gather, advantage moments, fused losses, clip+Adam)
* `rollout`  the time-major device rollout buffer the path consumes
* `agents`   drop-in `PPO` / `A2C` classes with the reference's constructors and method surface
* `dist`     env sharding across ranks + the two collectives (gradient all-reduce, advantage moments)

There is no CPU fallback anywhere: without the built library or without a CUDA device, calls raise.
"""
__version__ = '1.0.0'

from . import _ffi  # noqa: F401  (does not load the library until first use)


def library_path():
    return _ffi.library_path()
