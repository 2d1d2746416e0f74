"""xagents_b200 -- B200-native on-policy rollout-to-update hot path of abstractguy/xagents.

Only what that path needs lives here:

* `csrc/`    hand-written sm_100a CUDA kernels behind the C ABI of `include/xagents_b200.h`
* `_ffi`     ctypes binding of that ABI;  `_dlpack`  zero-copy device pointers from DLPack capsules
* `ops`      tensor-level calls (returns/GAE, permute-gather, advantage moments, fused losses, clip+Adam)
* `hotpath`  the prepared train-step pipelines (`PPOHotPath`, `A2CHotPath`): resolved launch tables, two streams, CUDA graph
* `agents`   drop-in `PPO` / `A2C` / `TRPO` / `ACER` classes with the reference's constructors and method surface,
             the `.cfg` network reader, the tensor-core (tcgen05) policy/value network
* `buffers`  ACER's replay storage as a device-resident trajectory ring read by the gather kernel
* `cli`      `python -m xagents_b200 train <agent> ...`: the reference's command line and factory for those agents
* `envs`     the built-in environments of BASELINE's configs (gym is absent from the image)
* `dist`     env sharding across ranks + the two collectives (gradient all-reduce, advantage moments)

There is no CPU fallback anywhere: without the built library or without a CUDA device, calls raise.
"""
__version__ = '1.0.0'

from . import _ffi  # noqa: F401  (does not load the library until first use)


def library_path():
    return _ffi.library_path()
