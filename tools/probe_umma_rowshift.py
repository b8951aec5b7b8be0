"""Hardware probe for DESIGN.md section 8 item 1(b): tcgen05.mma reading a SWIZZLE_128B K-major A operand from a start address
that is not 1024-byte aligned (row shift) and with a non-1024-byte stride between 8-row groups.  B is the identity, so the result
IS the A tile as the tensor core saw it; A[row][col] = row * 8 + (col >> 3) names the source row and 16-byte chunk of each element.
Writes gpurun_out/probe_umma_rowshift.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dim_b200 import _native
    ctx = _native.SelfTest(0)  # the probes live in libdimb200_selftest.so
    rows_a = 248
    A = (np.arange(rows_a)[:, None] * 8 + (np.arange(64)[None, :] >> 3)).astype(np.float32)
    B = np.eye(64, dtype=np.float32)
    C = np.zeros((128, 64), np.float32)
    out = []
    for sbo in (1024, 1152, 1280):
        for shift in (0, 1, 2, 3, 5, 8, 9):
            if shift + 15 * (sbo // 128) + 8 > rows_a:
                continue
            for bo in (0, 1):
                rc = ctx.lib.dimb_probe_rowshift(ctx.h, _native._ptr(A), _native._ptr(B), _native._ptr(C), rows_a, shift, sbo, bo)
                if rc != 0:
                    out.append({"sbo": sbo, "shift": shift, "base_offset": bo, "error": ctx.lib.dimb_last_error(ctx.h).decode()})
                    continue
                r = np.arange(128)
                src = shift + (r // 8) * (sbo // 128) + r % 8
                exp = src[:, None] * 8 + (np.arange(64)[None, :] >> 3)
                ok = bool(np.array_equal(C, exp))
                row_ok = bool(np.array_equal(C.astype(np.int64) >> 3, np.broadcast_to(src[:, None], C.shape)))
                rec = {"sbo": sbo, "shift": shift, "base_offset": bo, "matches_absolute_address_swizzle": ok, "rows_correct": row_ok}
                if not ok:  # describe what was read instead: (source row, source chunk) of the first elements of rows 0, 1, 8
                    rec["seen"] = {str(i): [[int(v) >> 3, int(v) & 7] for v in C[i, ::8]] for i in (0, 1, 8)}
                    rec["expected_rows"] = {str(i): int(src[i]) for i in (0, 1, 8)}
                out.append(rec)
                print(rec, flush=True)
    # ---- SWIZZLE_64B operands with 64-byte rows (half K block): A[row][col] = row * 4 + (col >> 3), B = I(32)
    A64 = (np.arange(rows_a)[:, None] * 4 + (np.arange(32)[None, :] >> 3)).astype(np.float32)
    B64 = np.eye(32, dtype=np.float32)
    C64 = np.zeros((128, 32), np.float32)
    for sbo in (512, 640):
        for shift in (0, 1, 2, 3, 5, 8, 10):
            if shift + 15 * (sbo // 64) + 8 > rows_a:
                continue
            rc = ctx.lib.dimb_probe_rowshift64(ctx.h, _native._ptr(A64), _native._ptr(B64), _native._ptr(C64), rows_a, shift, sbo)
            if rc != 0:
                out.append({"swizzle": "64B", "sbo": sbo, "shift": shift, "error": ctx.lib.dimb_last_error(ctx.h).decode()})
                print(out[-1], flush=True)
                break
            r = np.arange(128)
            src = shift + (r // 8) * (sbo // 64) + r % 8
            exp = src[:, None] * 4 + (np.arange(32)[None, :] >> 3)
            rec = {"swizzle": "64B", "sbo": sbo, "shift": shift, "matches_absolute_address_swizzle": bool(np.array_equal(C64, exp)),
                   "rows_correct": bool(np.array_equal(C64.astype(np.int64) >> 2, np.broadcast_to(src[:, None], C64.shape)))}
            if not rec["matches_absolute_address_swizzle"]:
                rec["seen"] = {str(i): [[int(v) >> 2, int(v) & 3] for v in C64[i, ::8]] for i in (0, 1, 8)}
                rec["expected_rows"] = {str(i): int(src[i]) for i in (0, 1, 8)}
            out.append(rec)
            print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_umma_rowshift.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
