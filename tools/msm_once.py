"""One warmed-up MSM of n points at a forced window width, for `ncu --metrics gpu__time_duration.sum` launch lists:
    B2P_MSM_C=19 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/msm_once.py BN254 131075"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from algoplonk_b200 import _lib, api  # noqa: E402

curve, n = sys.argv[1], int(sys.argv[2])
_lib.init(0)
lib = _lib.load()
srs = api.SRS.unsafe(curve, n)
gen = torch.Generator(device="cpu").manual_seed(1)
raw = torch.randint(-(1 << 31), 1 << 31, (n, 8), generator=gen, dtype=torch.int64).to(torch.int32)
raw[:, 7] &= 0x0FFFFFFF
d = raw.to("cuda:0").contiguous()
out = C.create_string_buffer(2 * api.FP_BYTES[curve])
for _ in range(int(os.environ.get("REPS", "2"))):
    _lib.check(lib.b2p_msm_g1_dev(srs.handle, _lib.BASIS_CANONICAL, d.data_ptr(), n, out))
print(srs.msm_params())
