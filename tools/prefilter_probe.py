"""Prefilter passes on the BASELINE grid sizes: builds a radial RIF of each size on the device twice — with the tiled x pass
(k_prefilter_x) and with the one-thread-per-line x pass (MER_PREFILTER_X_UNTILED=1) — checks that the results are
bit-identical and prints wall times of the whole handle creation.  Run under
`ncu --metrics gpu__time_duration.sum -k regex:k_prefilter --csv` for the per-pass times (profiles/r02_prefilter_passes.csv).
Usage: python tools/prefilter_probe.py [sizes...]   (default 256 512 1024)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mitsubaer_b200 as mer  # noqa: E402
from mitsubaer_b200._abi import lib  # noqa: E402

BOX_MIN, BOX_MAX = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)


def build(n, untiled):
    if untiled:
        os.environ["MER_PREFILTER_X_UNTILED"] = "1"
    else:
        os.environ.pop("MER_PREFILTER_X_UNTILED", None)
    res = (n, n, n)
    lo, hi = mer.fields.padded_bbox(BOX_MIN, BOX_MAX, res)
    data = mer.fields.radial_rif(res, lo, hi, xp=torch, device="cuda:0")
    torch.cuda.synchronize()
    t0 = time.time()
    rif = mer.SplineDataSource(data_ptr=data.data_ptr(), res=res, min=lo, max=hi, device=0, mode="tricubic")
    torch.cuda.synchronize()
    dt = time.time() - t0
    del data
    return rif, dt


def fingerprint(rif, n):
    if n <= 512:
        return rif.coefficients()  # every prefiltered coefficient, read back through the C ABI
    # 1024^3: the spline at 2^20 random points instead of copying 4 GiB to the host
    g = torch.Generator().manual_seed(5)
    p = (torch.rand((1 << 20, 3), generator=g) * 1.8 - 0.9).numpy()
    return rif.value(p)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
    build(64, False)  # context, pool
    for n in sizes:
        a, ta = build(n, False)
        fa = fingerprint(a, n)
        del a
        lib.mer_trim_memory(0)
        b, tb = build(n, True)
        fb = fingerprint(b, n)
        del b
        lib.mer_trim_memory(0)
        print(json.dumps({"grid": "%d^3" % n, "create_s_tiled_x": round(ta, 4), "create_s_untiled_x": round(tb, 4),
                          "bit_identical": bool((fa == fb).all()), "compared": int(fa.size)}), flush=True)


if __name__ == "__main__":
    main()
