"""A/B two builds of libfo1.so on the same box: time fo1_gemm_bf16 on a list of (M, N, K, act) shapes through each
library (ctypes, CUDA-event timing, L2 flushed).  Usage: python scripts/ab_gemm.py <libA.so> <libB.so>"""
import ctypes as C
import json
import sys

import numpy as np
import torch


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("A", C.c_void_p), ("lda", C.c_int64), ("W", C.c_void_p), ("ldw", C.c_int64),
                ("D", C.c_void_p), ("ldd", C.c_int64), ("d_dtype", C.c_int32),
                ("bias", C.c_void_p), ("bias_dtype", C.c_int32), ("act", C.c_int32),
                ("residual", C.c_void_p), ("ldr", C.c_int64), ("gated", C.c_int32), ("tile_n", C.c_int32), ("split_k", C.c_int32)]


SHAPES = [(100352, 4096, 1024, 1, 0, 0), (131072, 3840, 1280, 0, 0, 0), (131072, 1280, 1280, 0, 0, 1), (38240, 2048, 2048, 0, 0, 1),
          (38240, 2048, 11008, 0, 0, 1), (131072, 1280, 3424, 0, 0, 1), (131072, 6848, 1280, 2, 1, 0), (100352, 1024, 4096, 0, 0, 1)]


def main():
    libs = [(p, C.CDLL(p)) for p in sys.argv[1:]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, N, K, act, gated, resid in SHAPES:
        a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N // 2 if gated else N, device="cuda", dtype=torch.bfloat16)
        d = GemmDesc()
        d.M, d.N, d.K = M, N, K
        d.A, d.lda, d.W, d.ldw, d.D, d.ldd, d.d_dtype = a.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), out.shape[1], 0
        d.bias, d.bias_dtype, d.act, d.gated = bias.data_ptr(), 0, act, gated
        if resid:
            res = torch.randn_like(out)
            d.residual, d.ldr = res.data_ptr(), out.shape[1]
        row = {"M": M, "N": N, "K": K, "act": act, "gated": gated, "resid": resid}
        for path, lib in libs:
            lib.fo1_gemm_bf16.restype = C.c_int
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            ts = []
            for i in range(8):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.fo1_gemm_bf16(C.byref(d), st)
                e1.record(); torch.cuda.synchronize()
                assert rc == 0, rc
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            row[path.split("/")[-1]] = {"ms": round(ms, 3), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
        print(json.dumps(row), flush=True)
        del a, w, out


if __name__ == "__main__":
    main()
