"""Time the MSS stage of the device-resident step for several scan chunk sizes (config 2)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepgrp_b200 import _lib, model
L = int(sys.argv[1]) if len(sys.argv) > 1 else 46_700_000
ctx = _lib.context(0)
w = model.random_weights(342, 60, attention=True, seed=0)
h = w.device_handle(ctx)
codes = torch.from_numpy(np.random.default_rng([1, 0]).integers(0, 4, size=L, dtype=np.uint8)).cuda()
n_rows = ctypes.c_int64(0)
for ch in (0, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
    ctx.set_int("mss_chunk", ch)
    for rep in range(2):
        _lib.check(_lib.lib().dgrp_predict_codes_dev(ctx.handle, h, ctypes.c_void_p(codes.data_ptr()), L, 50, 256, 1,
                                                      50, 50, 0, ctypes.byref(n_rows)))
    t = ctx.timings()
    print("chunk %6d  mss_ms %.3f  rounds %d  forward_ms %.1f" % (ch, t["mss_ms"], ctx.get_int("mss_rounds"), t["forward_ms"]), flush=True)
