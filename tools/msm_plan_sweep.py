"""MSM time against (points, window bits): the data behind msm_plan's choices (csrc/msm_digits.cuh).  For every
n in the list an unsafe SRS of n points is generated with each forced window width c (B2P_MSM_C) and one n-scalar
MSM (uniform scalars resident in HBM, b2p_msm_g1_dev: sort + accumulate + full reduction + host inversion) is
timed with CUDA events on the SRS handle's stream.  One JSON line per (curve, n, c).

    python tools/msm_plan_sweep.py BN254 17,18,19,20 15,16,17,18,19,20
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from algoplonk_b200 import _lib, api  # noqa: E402


def main():
    curve = sys.argv[1] if len(sys.argv) > 1 else "BN254"
    logs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "17,18,19,20").split(",")]
    cs = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,15,16,17,18,19,20").split(",")]
    _lib.init(0)
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    for lg in logs:
        n = (1 << lg) + 3
        gen = torch.Generator(device="cpu").manual_seed(lg)
        raw = torch.randint(-(1 << 31), 1 << 31, (n, 8), generator=gen, dtype=torch.int64).to(torch.int32)
        raw[:, 7] &= 0x0FFFFFFF
        d_scalars = raw.to(dev).contiguous()
        for c in cs:
            if c:
                os.environ["B2P_MSM_C"] = str(c)
            else:
                os.environ.pop("B2P_MSM_C", None)
            srs = api.SRS.unsafe(curve, n)
            cc, W, nb = srs.msm_params()
            out = C.create_string_buffer(2 * api.FP_BYTES[curve])
            stream = torch.cuda.ExternalStream(lib.b2p_srs_stream(srs.handle), device=dev)
            for _ in range(3):
                _lib.check(lib.b2p_msm_g1_dev(srs.handle, _lib.BASIS_CANONICAL, d_scalars.data_ptr(), n, out))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                _lib.check(lib.b2p_msm_g1_dev(srs.handle, _lib.BASIS_CANONICAL, d_scalars.data_ptr(), n, out))
            e1.record(stream)
            e1.synchronize()
            print(json.dumps({"curve": curve, "points": n, "forced_c": c, "c": cc, "windows": W, "buckets": nb,
                              "ms_per_msm": e0.elapsed_time(e1) / 10}), flush=True)
            srs.free()
    os.environ.pop("B2P_MSM_C", None)


if __name__ == "__main__":
    main()
