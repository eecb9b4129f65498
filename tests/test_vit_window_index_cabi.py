"""CPU: the engine's integer bookkeeping (window permutation, cu_seqlens, patch coordinates), called
through the C ABI, is bit-exact against the reference's get_window_index / rot_pos_emb goldens."""
import ctypes as C

import numpy as np

from tests.golden_io import load


def test_window_index_bit_exact_through_cabi():
    import fo1_b200
    from importlib import import_module
    E = import_module("vlm-fo1_b200.engine")
    from oracle import vit as OV
    L = fo1_b200.lib()
    cfg = E.EngineConfig().to_c()
    L.fo1_vit_window_index.restype = C.c_int
    _, z = load("vit_small")
    for key in [k for k in z if k.startswith("wi_")]:
        gh, gw = (int(v) for v in key[3:].split("x"))
        n = gh * gw
        wi = (C.c_int32 * (n // 4))(); cu = (C.c_int32 * (n // 4 + 2))(); ncu = C.c_int32(); pos = (C.c_int32 * (2 * n))()
        rc = L.fo1_vit_window_index(C.byref(cfg), gh, gw, wi, cu, C.byref(ncu), pos)
        assert rc == 0, L.fo1_last_error()
        assert np.array_equal(np.frombuffer(wi, dtype=np.int32), z[key]), key
        assert np.array_equal(np.frombuffer(cu, dtype=np.int32)[: ncu.value], z["cu_" + key[3:]]), key
        # patch coordinates in window order == rot_pos_emb positions permuted by the window index
        ref_pos = OV.patch_positions(gh, gw).reshape(n // 4, 4, 2)[z[key].astype(np.int64)].reshape(n, 2).numpy()
        assert np.array_equal(np.frombuffer(pos, dtype=np.int32).reshape(n, 2), ref_pos), key
    assert L.fo1_vit_window_index(C.byref(cfg), 5, 4, wi, cu, C.byref(ncu), pos) != 0   # odd grid rejected
