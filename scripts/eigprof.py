"""Per-phase timing of the device eigensolver (SM clock at the phase boundaries, bsub_debug_eig_cycles)."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import background_subtraction_b200 as B  # noqa: E402
from background_subtraction_b200 import _cabi as C, synth  # noqa: E402

for (rows, cols, n) in [(240, 320, 48), (240, 320, 200), (240, 320, 300), (120, 160, 600)]:
    video, _ = synth.make_clip(rows, cols, n, seed=1, n_rect=3)
    D = synth.preprocess_u8(video)
    cfg = B.make_config(rows * cols, n, C.PROX_FLAT_LINF, rows, cols)
    dec = B.Decomposition(cfg)
    dec.set_flat_groups(B.get_proximal_flat_groups_nonoverlap((rows, cols), (3, 3)))
    dec.load(D)
    dec.run()
    st = dec.status()
    out = (ctypes.c_int64 * 16)()
    C.check(dec.lib.bsub_debug_eig_cycles(dec.h, out))
    d = np.diff(np.array(list(out), dtype=np.float64)[:6]) / 1.965e3
    log = dec.log()
    print(f"n={n} iters={st.iter} sv_last={log[-1]['sv']} svp={log[-1]['svp']} us: tridiag={d[0]:.0f} eigval={d[1]:.0f} "
          f"invit={d[2]:.0f} reorth={d[3]:.0f} backtr={d[4]:.0f} total={np.sum(d):.0f}  tridiag sub-phases (us): "
          + " ".join(f"{k}={v / 1.965e3:.0f}" for k, v in zip(["reflector", "matvec", "sync2", "dot_w", "update", "sync1"], list(out)[8:14])))
