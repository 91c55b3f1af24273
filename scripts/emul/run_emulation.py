"""Runs scripts/emul/graph_tile_emulation.c on the input of the FIRST prox call of a WaterSurface graph-LSD solve (lambda/mu = 8.5e-3,
the hardest one: the oracle's sequential sweeps need ~3600 sweeps at tol = 1e-6 lambda) and compares with the oracle at 1e-13.
Test infrastructure (imports oracle/); CPU only.   python scripts/emul/run_emulation.py"""
import os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import alm_oracle as O                      # noqa: E402
import background_subtraction_b200 as B                 # noqa: E402  (host-side graph builder only)


class _Stop(Exception):
    pass


def main():
    exe = os.path.join(tempfile.gettempdir(), "graph_tile_emulation")
    subprocess.check_call(["gcc", "-O2", "-DNOISE=2.4e-7f", "-o", exe, os.path.join(HERE, "graph_tile_emulation.c"), "-lm"])
    cube = np.asfortranarray(np.load(os.path.join(ROOT, "tests", "golden", "watersurface_u8.npz"))["ImData"])
    D, _x, _m = O.normalize_and_center(cube[:, :, :16])
    rows, cols = 128, 160
    gc = O.graph_from_spams_dict(B.getGraphSPAMS_all_groups((rows, cols), (3, 3)))
    got = {}

    def prox_fn(G_S, lam, mu):
        got["u"], got["lam1"] = G_S[:, [3]].copy(order='F'), float(lam / mu)
        raise _Stop()
    try:
        O._alm(D, prox_fn, 12.5, 10)
    except _Stop:
        pass
    u, lam1 = got["u"], got["lam1"]
    ref, sw = O.prox_graph(u, lam1, gc, tol=1e-13, max_sweeps=100000, return_sweeps=True)
    _r6, sw6 = O.prox_graph(u, lam1, gc, tol=1e-6 * lam1, max_sweeps=100000, return_sweeps=True)
    with tempfile.TemporaryDirectory() as td:
        up, vp = os.path.join(td, "u.bin"), os.path.join(td, "v.bin")
        u[:, 0].astype(np.float32).tofile(up)
        r = subprocess.run([exe, str(rows), str(cols), repr(lam1), repr(1e-6 * lam1), "4000", up, vp], capture_output=True, text=True)
        v = np.fromfile(vp, dtype=np.float32)
    err = np.abs(v - ref[:, 0])
    print("lambda/mu %.3e: oracle sweeps %d (tol 1e-13) / %d (tol 1e-6 lambda); tiled schedule: %s; max |x - x_oracle| %.2e, relF %.2e"
          % (lam1, int(sw[0]), int(sw6[0]), r.stderr.strip().splitlines()[-1], err.max(), np.linalg.norm(err) / np.linalg.norm(ref)))


if __name__ == "__main__":
    main()
