"""GPU micro-benchmark: every GEMV variant (WK,RP,U) on the llama2-7B / stories110M matrix shapes.
Prints achieved GB/s (weight bytes / CUDA-event time per launch) against the measured HBM peak.
Working sets > L2 are cycled through several copies so no launch re-reads a cached matrix."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rama_b200 import _lib  # noqa: E402
from rama_b200.engine import GPU, DeviceBuffer  # noqa: E402

VARIANTS = {0: "WK16 RP2 U2", 1: "WK8 RP2 U4", 2: "WK4 RP2 U4", 3: "WK1 RP2 U4", 4: "WK16 RP4 U2",
            5: "WK8 RP4 U2", 6: "WK2 RP2 U4", 7: "WK16 RP1 U4"}
SHAPES = [("7B wq/wo", 4096, 4096), ("7B qkv-like", 12288, 4096), ("7B w1w3-like", 22016, 4096),
          ("7B w2", 4096, 11008), ("7B wcls", 32000, 4096), ("110M qkv-like", 2304, 768), ("110M w2", 768, 2048)]


def main():
    gpu = GPU(0)
    peak = 6528.7
    out = []
    for name, rows, width in SHAPES:
        nbytes = rows * width * 4
        copies = max(1, min(64, -(-600_000_000 // nbytes)))  # cycle ≥ 600 MB (> L2) of distinct matrices
        w = DeviceBuffer(gpu, rows * width * copies)
        _lib.check(_lib.lib().rama_synth_fill(gpu.h, w.ptr(), w.n, 1, 2, 0, 0.01, 0.0))
        x = DeviceBuffer(gpu, width)
        _lib.check(_lib.lib().rama_synth_fill(gpu.h, x.ptr(), x.n, 2, 3, 0, 1e-5, 0.0))
        o = DeviceBuffer(gpu, rows)
        for v, vn in VARIANTS.items():
            ms = C.c_float()
            _lib.check(_lib.lib().rama_bench_gemv(gpu.h, o.ptr(), w.ptr(0), x.ptr(), rows, width, copies, v,
                                                  10 * copies, C.byref(ms)))
            ms_avg = ms.value
            gbs = nbytes / (ms_avg * 1e-3) / 1e9
            out.append({"shape": name, "rows": rows, "width": width, "variant": v, "cfg": vn,
                        "us": round(ms_avg * 1e3, 2), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3),
                        "copies": copies})
            print(f"{name:16s} {rows:6d}x{width:<6d} v{v} {vn:12s} {ms_avg*1e3:9.2f} us {gbs:8.1f} GB/s {gbs/peak:6.3f}", flush=True)
        w.free(); x.free(); o.free()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemv_sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()
