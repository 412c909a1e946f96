import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genedex_b200 as gdx
lib = gdx._lib.load()
for rec in (32, 64, 128):
    g, l = C.c_double(), C.c_double()
    assert lib.gdx_measure_random_gather(0, 12 << 30, rec, 1 << 28, 0, C.byref(g), C.byref(l)) == 0
    print(rec, g.value, l.value)
