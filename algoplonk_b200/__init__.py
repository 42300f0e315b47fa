"""algoplonk_b200: B200-native PLONK prover behind AlgoPlonk's proving path.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI), the
ctypes binding, a minimal constraint-system front-end for tests/benchmarks and
the host-side mirror of AlgoPlonk's Compile / Verify / MarshalProof.
"""
from . import frontend  # noqa: F401
from ._lib import B200PlonkError, LIB_PATH  # noqa: F401
