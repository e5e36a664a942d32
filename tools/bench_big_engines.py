import sys, json
sys.path.insert(0, 'prgs-sdr-kspecanal_b200')
import numpy as np
from kspec import _ffi, synth
from kspec.engine import Plan
from kspec.hotpath import derive_config
for F, win, r, prec, n in ((16384, "ones", 0.1, "auto", 32), (16384, "ones", 0.1, "f32", 32), (32768, "hanning", 0.5, "auto", 16), (1 << 18, "ones", 0.1, "auto", 4), (1200, "hanning", 0.1, "auto", 256), (100000, "ones", 0.5, "auto", 4)):
    d = derive_config(dict(fftSize=F, window=win, samplingRate=2.4e6))
    S = d["fullSize"]
    x = synth.tones_noise(n * S, seed=1)
    plan = Plan(F, S, r, d["theWin"], "AVG", _ffi.IN_C64, precision=prec)
    dp = plan.dev_alloc(n * S * 8); plan.dev_upload(dp, x)
    for _ in range(2): plan.zerospan_batch_dev(dp, n, 19.1, d["xRes"], "MAX")
    plan.sync(); plan.timer_start()
    for _ in range(3): plan.zerospan_batch_dev(dp, n, 19.1, d["xRes"], "MAX")
    ms = plan.timer_stop() / 3
    print(json.dumps(dict(F=F, r=r, prec=plan.precision, path=plan.path, frames=plan.n_frames, scans=n, ms=round(ms, 3), MSps=round(n * S / ms / 1e3, 1), launches=plan.launch_count())))
    plan.dev_free(dp); plan.close()
