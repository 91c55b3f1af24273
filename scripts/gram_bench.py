"""Times the int8 Gram kernel alone (profiling hook bsub_gram_i8_bench).  BSUB_GRAM_DBG selects the profiling knobs of
gram_i8_c3_kernel (low 4 bits: digit pairs issued per k-step, 0x100: no loads) -- one process per setting because the knob is
read once."""
import ctypes, os, subprocess, sys

def one(n, ldq, reps=20):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from background_subtraction_b200 import _cabi as C
    ms = ctypes.c_float(0)
    C.check(C.load().bsub_gram_i8_bench(n, ldq, reps, ctypes.byref(ms)))
    return ms.value

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        print("n=%s ldq=%s dbg=%s: %.4f ms" % (sys.argv[2], sys.argv[3], os.environ.get("BSUB_GRAM_DBG", "0"), one(int(sys.argv[2]), int(sys.argv[3]))), flush=True)
    else:
        shapes = [(300, 2073600), (300, 259200)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
        for n, ldq in shapes:
            for dbg in os.environ.get("GRAM_DBG_LIST", "0,1,5,0x100,0x101,0x200,0x201,0x300,0x301").split(","):
                env = dict(os.environ, BSUB_GRAM_DBG=dbg)
                subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(n), str(ldq)], env=env, check=False)
